// RDS bit path shared by the host decoder (rds_host.cpp) and the device kernel K6 (k6_rds.cu):
// one source, so both are the same integer program and bit-identical by construction.
//   soft BPSK symbols -> differential-Manchester bits -> 16-byte packets -> 26-bit blocks
//   -> groups -> PI / PTY / PS / RadioText
// Behaviour follows, bit for bit (reference file:line under /root/reference/src):
//   DifferentialManchesterDecoder::Process/PushBit  rds_decoder/differential_manchester_decoder.h:25-59
//   RDS_Group_Sync (FindingSync / ReadingGroup)     rds_decoder/rds_group_sync.cpp:29-237
//   CalculateCRC10 / single-bit error patterns      rds_decoder/crc10.cpp:9-60, rds_constants.h:15-28
//   RDS_Decoder::ProcessGroup, groups 0A 2A 4A 10A  rds_decoder/rds_decoder.cpp:82-126, 159-245, 301-340, 363-441
//   RDS_Database_Decoder_Handler                    rds_decoder/rds_database_decoder_handler.cpp:15-138
//   mjd_to_ymd                                      rds_decoder/modified_julian_date.h:8-24
#pragma once
#include <stdint.h>
#include "../../include/fmgpu.h"

#if defined(__CUDACC__)
#define RDS_HD __host__ __device__ __forceinline__
#else
#define RDS_HD inline
#endif

namespace rds {

constexpr uint32_t CRC10_POLY = 0b0110111001;                    // x^10 implicit (rds_constants.h:15)
enum { OFF_A = 0, OFF_B, OFF_C, OFF_C1, OFF_D };

RDS_HD uint32_t offset_word(int id) {                            // rds_constants.h:20-28
    switch (id) {
    case OFF_A: return 0b0011111100;
    case OFF_B: return 0b0110011000;
    case OFF_C: return 0b0101101000;
    case OFF_C1: return 0b1101010000;
    case OFF_D: return 0b0110110100;
    default: return 0;
    }
}

RDS_HD uint32_t syndrome_of(uint32_t codeword) {                 // crc10.cpp:9-25 over 26 bits
    uint32_t reg = 0;
    for (int bit = 25; bit >= 0; bit--) {
        reg = (reg << 1) | ((codeword >> bit) & 1u);
        if (reg & 0x400u) reg ^= CRC10_POLY;
    }
    return reg & 0x3FFu;
}

// The syndrome is linear over GF(2), so it is the XOR of per-byte table entries (4 lookups instead of
// a 26-step shift register), and a non-zero syndrome names at most one single-bit error position
// (the reference builds that map in crc10.cpp:28-52; single-bit syndromes of this code are distinct).
// Built once on the host (build_tables); the device kernel stages a copy in shared memory.
struct Tables {
    uint16_t syn[4][256];       // syn[j][v] = syndrome_of(v << 8j)
    uint8_t errpos[1024];       // syndrome -> bit index + 1 of the single-bit error, 0 = none
};

inline void build_tables(Tables& T) {
    for (int j = 0; j < 4; j++)
        for (uint32_t v = 0; v < 256; v++) {
            const uint32_t w = (v << (8 * j)) & 0x3FFFFFFu;
            T.syn[j][v] = (uint16_t)syndrome_of(w);
        }
    for (int i = 0; i < 1024; i++) T.errpos[i] = 0;
    for (int i = 0; i < 26; i++) T.errpos[syndrome_of(1u << i)] = (uint8_t)(i + 1);
}

RDS_HD uint32_t syndrome_fast(const Tables& T, uint32_t x) {
    return (uint32_t)(T.syn[0][x & 255u] ^ T.syn[1][(x >> 8) & 255u] ^ T.syn[2][(x >> 16) & 255u] ^ T.syn[3][(x >> 24) & 3u]);
}

RDS_HD uint32_t single_bit_pattern(const Tables& T, uint32_t syn) {
    const uint32_t p = T.errpos[syn & 1023u];
    return p ? (1u << (p - 1u)) : 0u;
}

// Per-stream decoder state as it is stored (host object / device array): plain data.
struct State {
    uint32_t pk[4];                    // the 16-byte packet being filled; byte b = bits 8(b&3).. of pk[b>>2]
    uint32_t shift;                    // last 26 bits
    uint8_t byte_index, bit_index, take, prev_level;
    uint8_t locked, rt_ab, block_slot, block_errors;
    uint8_t bits_in_block, bad_groups, pty, pad0;
    uint16_t pi, pad1;
    fmgpu_rds_group cur;               // group being assembled
    char ps[8];
    char rt[64];
    unsigned long long n_groups, n_bytes;      // totals since creation
    fmgpu_rds_db_ext ext;              // flags, TA, PTYN, clock: written a few times a second
};

// The same state while a decoder runs: scalars only (no indexed arrays), so that on the device it
// lives in registers -- with the stored layout used directly the whole struct went to local memory
// and the kernel took 0.08 ms alone / 0.18 ms next to K3 (measured).  The RadioText buffer, touched
// once per 2A group, and the rest of the database (0A flags, 4A clock, 10A name) stay in memory
// behind `rt` / `ext`.
struct Work {
    uint32_t pk0, pk1, pk2, pk3, shift;
    uint32_t byte_index, bit_index, take, prev_level, locked, rt_ab, block_slot, block_errors, bits_in_block, bad_groups, pty, pi;
    unsigned long long cur_data;       // data[slot] at bits 16 slot
    uint32_t cur_valid, cur_type;      // valid[slot], type[slot] at bits 8 slot
    unsigned long long ps;             // ps[i] at bits 8 i
    unsigned long long n_groups, n_bytes;
    char* rt;
    fmgpu_rds_db_ext* ext;
};

RDS_HD void init(State& s) {
    unsigned char* p = (unsigned char*)&s;
    for (unsigned i = 0; i < sizeof(State); i++) p[i] = 0;
    s.rt_ab = 0b100;                                             // "unknown" A/B flag: first 2A group clears the text
    s.ext.ptyn_ab_flag = 0b100;                                  // same for the programme type name (handler.h:12)
}

RDS_HD void load(Work& w, State& s) {
    w.pk0 = s.pk[0]; w.pk1 = s.pk[1]; w.pk2 = s.pk[2]; w.pk3 = s.pk[3]; w.shift = s.shift;
    w.byte_index = s.byte_index; w.bit_index = s.bit_index; w.take = s.take; w.prev_level = s.prev_level;
    w.locked = s.locked; w.rt_ab = s.rt_ab; w.block_slot = s.block_slot; w.block_errors = s.block_errors;
    w.bits_in_block = s.bits_in_block; w.bad_groups = s.bad_groups; w.pty = s.pty; w.pi = s.pi;
    w.cur_data = 0; w.cur_valid = 0; w.cur_type = 0; w.ps = 0;
    for (int i = 0; i < 4; i++) {
        w.cur_data |= (unsigned long long)s.cur.data[i] << (16 * i);
        w.cur_valid |= (uint32_t)s.cur.valid[i] << (8 * i);
        w.cur_type |= (uint32_t)s.cur.type[i] << (8 * i);
    }
    for (int i = 0; i < 8; i++) w.ps |= (unsigned long long)(unsigned char)s.ps[i] << (8 * i);
    w.n_groups = s.n_groups; w.n_bytes = s.n_bytes;
    w.rt = s.rt;
    w.ext = &s.ext;
}

RDS_HD void group_of(const Work& w, fmgpu_rds_group& g) {
    for (int i = 0; i < 4; i++) {
        g.data[i] = (uint16_t)(w.cur_data >> (16 * i));
        g.valid[i] = (uint8_t)(w.cur_valid >> (8 * i));
        g.type[i] = (uint8_t)(w.cur_type >> (8 * i));
    }
}

RDS_HD void store(const Work& w, State& s) {
    s.pk[0] = w.pk0; s.pk[1] = w.pk1; s.pk[2] = w.pk2; s.pk[3] = w.pk3; s.shift = w.shift;
    s.byte_index = (uint8_t)w.byte_index; s.bit_index = (uint8_t)w.bit_index; s.take = (uint8_t)w.take; s.prev_level = (uint8_t)w.prev_level;
    s.locked = (uint8_t)w.locked; s.rt_ab = (uint8_t)w.rt_ab; s.block_slot = (uint8_t)w.block_slot; s.block_errors = (uint8_t)w.block_errors;
    s.bits_in_block = (uint8_t)w.bits_in_block; s.bad_groups = (uint8_t)w.bad_groups; s.pty = (uint8_t)w.pty; s.pi = (uint16_t)w.pi;
    group_of(w, s.cur);
    for (int i = 0; i < 8; i++) s.ps[i] = (char)(w.ps >> (8 * i));
    s.n_groups = w.n_groups; s.n_bytes = w.n_bytes;
}

RDS_HD bool try_block(Work& w, const Tables& T, uint32_t raw, uint32_t offset_id, uint32_t slot) {
    const uint32_t x = raw ^ offset_word((int)offset_id);
    uint32_t fixed = x;
    bool ok = false;
    const uint32_t syn = syndrome_fast(T, x);
    if (syn == 0) ok = true;
    else {
        const uint32_t e = single_bit_pattern(T, syn);
        if (e != 0 && syndrome_fast(T, x ^ e) == 0) { fixed = x ^ e; ok = true; }
    }
    w.cur_type = (w.cur_type & ~(0xFFu << (8 * slot))) | (offset_id << (8 * slot));
    w.cur_data = (w.cur_data & ~(0xFFFFull << (16 * slot))) | ((unsigned long long)((fixed >> 10) & 0xFFFFu) << (16 * slot));
    w.cur_valid = (w.cur_valid & ~(0xFFu << (8 * slot))) | ((ok ? 1u : 0u) << (8 * slot));
    return ok;
}

RDS_HD void push_block(Work& w, const Tables& T, uint32_t raw) {
    const uint32_t slot = w.block_slot;
    // offset word expected in this slot (A, B, C then C', D): rds_group_sync.cpp:150-190
    const uint32_t first = slot == 0 ? OFF_A : slot == 1 ? OFF_B : slot == 2 ? OFF_C : OFF_D;
    bool ok = try_block(w, T, raw, first, slot);
    if (!ok && slot == 2) ok = try_block(w, T, raw, OFF_C1, slot);
    w.block_slot++;
    if (!ok) w.block_errors++;
}

RDS_HD unsigned long long put_char64(unsigned long long v, int idx, uint32_t c) {
    const unsigned long long ch = (c == '\r') ? 0ull : (unsigned long long)(c & 0xFFu);
    return (v & ~(0xFFull << (8 * idx))) | (ch << (8 * idx));
}
RDS_HD void put_char(char* dst, int idx, uint32_t c) { dst[idx] = (c == '\r') ? 0 : (char)c; }

RDS_HD void update_database(Work& w) {
    const uint32_t d0 = (uint32_t)(w.cur_data & 0xFFFFu), bw = (uint32_t)((w.cur_data >> 16) & 0xFFFFu);
    const uint32_t d2 = (uint32_t)((w.cur_data >> 32) & 0xFFFFu), d3 = (uint32_t)((w.cur_data >> 48) & 0xFFFFu);
    const uint32_t v0 = w.cur_valid & 0xFFu, v1 = (w.cur_valid >> 8) & 0xFFu, v2 = (w.cur_valid >> 16) & 0xFFu, v3 = (w.cur_valid >> 24) & 0xFFu;
    if (v0) w.pi = d0;
    if (!v1) return;
    w.pty = (bw >> 5) & 31u;
    if ((bw >> 11) & 1u) return;                                 // version B: unsupported in the reference
    const uint32_t code = bw >> 12;
    const bool has_c = v2 && ((w.cur_type >> 16) & 0xFFu) == OFF_C;
    const bool has_d = v3 && ((w.cur_type >> 24) & 0xFFu) == OFF_D;
    if (code == 0) {
        const int seg = (int)(bw & 3u);
        fmgpu_rds_db_ext& x = *w.ext;
        x.is_music = (uint8_t)((bw >> 3) & 1u);
        x.traffic_announcement = (uint8_t)((((bw >> 10) & 1u) << 1) | ((bw >> 4) & 1u));   // TP:TA, table 8
        const uint8_t di = (uint8_t)((bw >> 2) & 1u);              // decoder identification bit d(3 - seg), table 9
        if (seg == 0) x.is_dynamic_program_type = di;
        else if (seg == 1) x.is_compressed = di;
        else if (seg == 2) x.is_artificial_head = di;
        else x.is_stereo = di;
        if (has_d) { w.ps = put_char64(w.ps, 2 * seg, d3 >> 8); w.ps = put_char64(w.ps, 2 * seg + 1, d3 & 0xFFu); }
    } else if (code == 2) {
        const uint32_t ab = (bw >> 4) & 1u;
        const int seg = (int)(bw & 15u);
        if (ab != w.rt_ab) for (int i = 0; i < 64; i++) w.rt[i] = 0;
        w.rt_ab = ab;
        if (has_c) { put_char(w.rt, 4 * seg, d2 >> 8); put_char(w.rt, 4 * seg + 1, d2 & 0xFFu); }
        if (has_d) { put_char(w.rt, 4 * seg + 2, d3 >> 8); put_char(w.rt, 4 * seg + 3, d3 & 0xFFu); }
    } else if (code == 4) {
        fmgpu_rds_db_ext& x = *w.ext;
        if (has_c) {                                             // OnDate: Modified Julian Day -> calendar date
            int J = (int)(((bw & 3u) << 15) | (d2 >> 1)) + 2400001 + 68569;
            const int C = 4 * J / 146097;
            J = J - (146097 * C + 3) / 4;
            const int Y = 4000 * (J + 1) / 1461001;
            J = J - 1461 * Y / 4 + 31;
            const int M = 80 * J / 2447;
            x.day = (uint8_t)(J - 2447 * M / 80);
            J = M / 11;
            x.month = (uint8_t)(M + 2 - 12 * J);
            x.year = 100 * (C - 49) + Y + J;
        }
        if (has_c && has_d) { x.hour = (uint8_t)(((d2 & 1u) << 4) | (d3 >> 12)); x.minute = (uint8_t)((d3 >> 6) & 63u); }
        if (has_d) { const int v = (int)(d3 & 31u); x.local_time_offset = (int8_t)(((d3 >> 5) & 1u) ? -v : v); }
    } else if (code == 10) {
        fmgpu_rds_db_ext& x = *w.ext;
        const uint8_t ab = (uint8_t)((bw >> 4) & 1u);
        const int seg = (int)(bw & 1u);
        if (ab != x.ptyn_ab_flag) for (int i = 0; i < 8; i++) x.programme_type_name[i] = 0;
        x.ptyn_ab_flag = ab;
        if (has_c) { put_char(x.programme_type_name, 4 * seg, d2 >> 8); put_char(x.programme_type_name, 4 * seg + 1, d2 & 0xFFu); }
        if (has_d) { put_char(x.programme_type_name, 4 * seg + 2, d3 >> 8); put_char(x.programme_type_name, 4 * seg + 3, d3 & 0xFFu); }
    }
}

template <class Sink>
RDS_HD void push_bit(Work& w, const Tables& T, uint32_t bit, Sink& sink) {
    w.shift = ((w.shift << 1) | (bit & 1u)) & 0x3FFFFFFu;
    if (!w.locked) {                                             // FindingSync, rds_group_sync.cpp:46-74
        if (syndrome_fast(T, w.shift ^ offset_word(OFF_A)) != 0) return;
        w.locked = 1;
        w.bits_in_block = 0;
        push_block(w, T, w.shift);
        return;
    }
    if (++w.bits_in_block != 26) return;                         // ReadingGroup, :76-127
    w.bits_in_block = 0;
    push_block(w, T, w.shift);
    if (w.block_slot < 4) return;
    fmgpu_rds_group g;
    group_of(w, g);
    sink.group(g, w.n_groups);
    w.n_groups++;
    update_database(w);
    const uint32_t errors = w.block_errors;
    w.block_slot = 0;
    w.block_errors = 0;
    if (errors == 0) { w.bad_groups = 0; return; }
    if (++w.bad_groups >= 3) { w.locked = 0; w.bad_groups = 0; }
}

// One soft symbol -> (every other one) one differential bit into the 16-byte packet, MSB first.
// Returns true when the packet is full: the caller must then run process_packet.
RDS_HD bool push_symbol_level(Work& w, uint32_t level) {       // level = (symbol > 0)
    w.take ^= 1u;                                                // every other chip (:36-37)
    if (!w.take) return false;
    const uint32_t bit = level ^ w.prev_level;
    w.prev_level = level;
    const uint32_t b = w.byte_index, sh = 8u * (b & 3u);
    const uint32_t clr = (w.bit_index == 0) ? ~(0xFFu << sh) : 0xFFFFFFFFu;
    const uint32_t set = bit << (sh + 7u - w.bit_index);
    const uint32_t word = b >> 2;
    if (word == 0) w.pk0 = (w.pk0 & clr) | set;
    if (word == 1) w.pk1 = (w.pk1 & clr) | set;
    if (word == 2) w.pk2 = (w.pk2 & clr) | set;
    if (word == 3) w.pk3 = (w.pk3 & clr) | set;
    if (++w.bit_index == 8) { w.bit_index = 0; w.byte_index++; }
    if (w.byte_index != 16) return false;
    w.byte_index = 0;
    return true;
}

// The reference hands over 16-byte packets (:55-58): the group sync sees bits in bursts of 128.
template <class Sink>
RDS_HD void process_packet(Work& w, const Tables& T, Sink& sink) {
    const uint32_t pk[4] = { w.pk0, w.pk1, w.pk2, w.pk3 };
    sink.packet(pk, w.n_bytes);
    w.n_bytes += 16;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 4; q++) {
        const uint32_t word = pk[q];
        for (uint32_t j = 0; j < 32; j++)                        // byte j>>3 of the word, MSB of each byte first
            push_bit(w, T, (word >> (8u * (j >> 3) + 7u - (j & 7u))) & 1u, sink);
    }
}

template <class Sink>
RDS_HD void push_symbol(Work& w, const Tables& T, float v, Sink& sink) {
    if (push_symbol_level(w, v > 0.0f ? 1u : 0u)) process_packet(w, T, sink);
}

} // namespace rds
