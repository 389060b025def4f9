// Host-side RDS bit path of the product: soft BPSK symbols -> bits -> 26-bit blocks -> groups ->
// PI / PTY / PS / RadioText, for callers that take the symbols off the device themselves (the
// Broadcast_FM_Demod shim's OnRDSOut observers).  The batch path decodes on the device (k6_rds.cu);
// both run the one integer program of rds_core.h.  Behaviour follows, bit for bit:
//   DifferentialManchesterDecoder::PushBit      rds_decoder/differential_manchester_decoder.h:32-59
//   RDS_Group_Sync                              rds_decoder/rds_group_sync.cpp:29-237
//   CalculateCRC10 / single-bit error patterns  rds_decoder/crc10.cpp:9-60, rds_constants.h:15-28
//   RDS_Decoder::ProcessGroup, 0A 2A 4A 10A     rds_decoder/rds_decoder.cpp:82-126, 159-245, 301-340, 363-441
//   RDS_Database_Decoder_Handler                rds_decoder/rds_database_decoder_handler.cpp:15-138
#include <cstring>
#include <vector>
#include "rds_core.h"

namespace { const rds::Tables& tables() { static const rds::Tables T = [] { rds::Tables t; rds::build_tables(t); return t; }(); return T; } }

struct fmgpu_rds {
    rds::State st;
    std::vector<fmgpu_rds_group> groups;
    std::vector<uint8_t> bytes;
    fmgpu_rds() { rds::init(st); }
    // sink of rds::push_symbol
    void group(const fmgpu_rds_group& g, unsigned long long) { groups.push_back(g); }
    void packet(const uint32_t pk[4], unsigned long long) {
        for (int b = 0; b < 16; b++) bytes.push_back((uint8_t)(pk[b >> 2] >> (8 * (b & 3))));
    }
};

extern "C" {

/* used by fmgpu_create to upload the same tables to the device */
const void* fmgpu_rds_tables_(size_t* bytes) { if (bytes) *bytes = sizeof(rds::Tables); return &tables(); }

fmgpu_rds* fmgpu_rds_create(void) { return new fmgpu_rds(); }
void fmgpu_rds_destroy(fmgpu_rds* r) { delete r; }
void fmgpu_rds_push_symbols(fmgpu_rds* r, const float* sym, size_t n) {
    if (!r || !sym) return;
    const rds::Tables& T = tables();
    rds::Work w;
    rds::load(w, r->st);
    for (size_t i = 0; i < n; i++) rds::push_symbol(w, T, sym[i], *r);
    rds::store(w, r->st);
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
    if (pi) *pi = r->st.pi;
    if (pty) *pty = r->st.pty;
    if (ps8) std::memcpy(ps8, r->st.ps, 8);
    if (rt64) std::memcpy(rt64, r->st.rt, 64);
}
void fmgpu_rds_get_db_ext(const fmgpu_rds* r, fmgpu_rds_db_ext* out) {
    if (r && out) *out = r->st.ext;
}

} // extern "C"
