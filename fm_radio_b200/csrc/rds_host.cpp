// Host-side RDS bit path of the product: soft BPSK symbols -> bits -> 26-bit blocks -> groups ->
// PI / PTY / PS / RadioText.  Integer work at 2375 symbols/s per stream; it stays on the host in
// this round (SURVEY.md 8f rank 1 moves it to the device).  Behaviour follows, bit for bit:
//   DifferentialManchesterDecoder::PushBit      rds_decoder/differential_manchester_decoder.h:32-59
//   RDS_Group_Sync                              rds_decoder/rds_group_sync.cpp:29-237
//   CalculateCRC10 / single-bit error patterns  rds_decoder/crc10.cpp:9-60, rds_constants.h:15-28
//   RDS_Decoder::ProcessGroup, 0A and 2A        rds_decoder/rds_decoder.cpp:82-126, 167-245, 301-340
//   RDS_Database_Decoder_Handler                rds_decoder/rds_database_decoder_handler.cpp:15-50
#include <array>
#include <cstring>
#include <vector>
#include "../../include/fmgpu.h"

namespace {

constexpr uint16_t CRC10_POLY = 0b0110111001;                    // x^10 implicit (rds_constants.h:15)
constexpr uint16_t OFFSET_WORDS[6] = { 0b0011111100, 0b0110011000, 0b0101101000, 0b1101010000, 0b0110110100, 0 };
enum { OFF_A = 0, OFF_B, OFF_C, OFF_C1, OFF_D };

uint16_t syndrome_of(uint32_t codeword) {
    uint32_t reg = 0;
    for (int bit = 25; bit >= 0; bit--) {
        reg = (reg << 1) | ((codeword >> bit) & 1u);
        if (reg & 0x400u) reg ^= CRC10_POLY;
    }
    return (uint16_t)(reg & 0x3FFu);
}

// syndrome -> single-bit error pattern (0 = none); built once
struct SyndromeTable {
    std::array<uint32_t, 1024> pattern{};
    SyndromeTable() {
        // same insertion order as crc10.cpp:36-49 (data bits first, then checksum bits); single-bit
        // syndromes of a 26-bit cyclic code are distinct, so the order does not matter
        for (int i = 10; i < 26; i++) pattern[syndrome_of(1u << i)] = 1u << i;
        for (int i = 0; i < 10; i++) pattern[syndrome_of(1u << i)] = 1u << i;
    }
};
const SyndromeTable& table() { static SyndromeTable t; return t; }

} // namespace

struct fmgpu_rds {
    // manchester
    uint8_t packet[16]{};
    int byte_index = 0, bit_index = 0;
    bool take = false, prev_level = false;
    // sync
    uint32_t shift = 0;
    int bits_in_block = 0, block_slot = 0, block_errors = 0, bad_groups = 0;
    bool locked = false;
    fmgpu_rds_group cur{};
    // results
    std::vector<fmgpu_rds_group> groups;
    std::vector<uint8_t> bytes;
    uint16_t pi = 0; uint8_t pty = 0; char ps[8]{}; char rt[64]{}; uint8_t rt_ab = 0b100;

    bool try_block(uint32_t raw, int offset_id, int slot) {
        const uint32_t x = raw ^ OFFSET_WORDS[offset_id];
        uint32_t fixed = x;
        bool ok = false;
        const uint16_t syn = syndrome_of(x);
        if (syn == 0) ok = true;
        else if (const uint32_t e = table().pattern[syn]) {
            if (syndrome_of(x ^ e) == 0) { fixed = x ^ e; ok = true; }
        }
        cur.type[slot] = (uint8_t)offset_id;
        cur.data[slot] = (uint16_t)(fixed >> 10);
        cur.valid[slot] = ok ? 1 : 0;
        return ok;
    }
    void push_block(uint32_t raw) {
        const int slot = block_slot;
        cur.valid[slot] = 0;
        switch (slot) {
        case 0: try_block(raw, OFF_A, slot); break;
        case 1: try_block(raw, OFF_B, slot); break;
        case 2: if (!try_block(raw, OFF_C, slot)) try_block(raw, OFF_C1, slot); break;
        case 3: try_block(raw, OFF_D, slot); break;
        }
        block_slot++;
        if (!cur.valid[slot]) block_errors++;
    }
    void on_group() {
        groups.push_back(cur);
        const uint16_t bw = cur.data[1];
        if (cur.valid[0]) pi = cur.data[0];
        if (!cur.valid[1]) return;
        pty = (bw >> 5) & 31;
        if ((bw >> 11) & 1) return;                              // version B: unsupported in the reference
        const int code = bw >> 12;
        const bool has_c = cur.valid[2] && cur.type[2] == OFF_C;
        const bool has_d = cur.valid[3] && cur.type[3] == OFF_D;
        auto put = [](char* dst, int idx, uint8_t c) { dst[idx] = (c == '\r') ? 0 : (char)c; };
        if (code == 0) {
            const int seg = bw & 3;
            if (has_d) { put(ps, 2 * seg, cur.data[3] >> 8); put(ps, 2 * seg + 1, cur.data[3] & 0xFF); }
        } else if (code == 2) {
            const uint8_t ab = (bw >> 4) & 1;
            const int seg = bw & 15;
            if (ab != rt_ab) std::memset(rt, 0, sizeof(rt));
            rt_ab = ab;
            if (has_c) { put(rt, 4 * seg, cur.data[2] >> 8); put(rt, 4 * seg + 1, cur.data[2] & 0xFF); }
            if (has_d) { put(rt, 4 * seg + 2, cur.data[3] >> 8); put(rt, 4 * seg + 3, cur.data[3] & 0xFF); }
        }
    }
    void push_bit(int bit) {
        shift = ((shift << 1) | (uint32_t)(bit & 1)) & 0x3FFFFFFu;
        if (!locked) {                                            // FindingSync, rds_group_sync.cpp:46-74
            if (syndrome_of(shift ^ OFFSET_WORDS[OFF_A]) != 0) return;
            locked = true;
            bits_in_block = 0;
            push_block(shift);
            return;
        }
        if (++bits_in_block != 26) return;                        // ReadingGroup, :76-127
        bits_in_block = 0;
        push_block(shift);
        if (block_slot < 4) return;
        on_group();
        const int errors = block_errors;
        block_slot = 0;
        block_errors = 0;
        if (errors == 0) { bad_groups = 0; return; }
        if (++bad_groups >= 3) { locked = false; bad_groups = 0; }
    }
    void push_symbol(float v) {
        take = !take;                                             // every other chip
        if (!take) return;
        const bool level = v > 0.0f;
        const int bit = (level != prev_level) ? 1 : 0;
        prev_level = level;
        if (bit_index == 0) packet[byte_index] = 0;
        packet[byte_index] |= (uint8_t)(bit << (7 - bit_index));
        if (++bit_index == 8) { bit_index = 0; byte_index++; }
        if (byte_index == 16) {                                   // the reference hands over 16-byte packets
            byte_index = 0;
            bytes.insert(bytes.end(), packet, packet + 16);
            for (int i = 0; i < 128; i++) push_bit((packet[i >> 3] >> (7 - (i & 7))) & 1);
        }
    }
};

extern "C" {

fmgpu_rds* fmgpu_rds_create(void) { return new fmgpu_rds(); }
void fmgpu_rds_destroy(fmgpu_rds* r) { delete r; }
void fmgpu_rds_push_symbols(fmgpu_rds* r, const float* sym, size_t n) {
    if (!r || !sym) return;
    for (size_t i = 0; i < n; i++) r->push_symbol(sym[i]);
}
int fmgpu_rds_n_groups(const fmgpu_rds* r) { return r ? (int)r->groups.size() : 0; }
int fmgpu_rds_get_groups(const fmgpu_rds* r, fmgpu_rds_group* out, int max_groups) {
    if (!r || !out) return 0;
    const int n = (int)r->groups.size() < max_groups ? (int)r->groups.size() : max_groups;
    if (n > 0) std::memcpy(out, r->groups.data(), sizeof(fmgpu_rds_group) * n);
    return n;
}
int fmgpu_rds_n_bytes(const fmgpu_rds* r) { return r ? (int)r->bytes.size() : 0; }
int fmgpu_rds_get_bytes(const fmgpu_rds* r, uint8_t* out, int max_bytes) {
    if (!r || !out) return 0;
    const int n = (int)r->bytes.size() < max_bytes ? (int)r->bytes.size() : max_bytes;
    if (n > 0) std::memcpy(out, r->bytes.data(), n);
    return n;
}
void fmgpu_rds_get_db(const fmgpu_rds* r, uint16_t* pi, char ps8[8], char rt64[64], uint8_t* pty) {
    if (!r) return;
    if (pi) *pi = r->pi;
    if (pty) *pty = r->pty;
    if (ps8) std::memcpy(ps8, r->ps, 8);
    if (rt64) std::memcpy(rt64, r->rt, 64);
}

} // extern "C"
