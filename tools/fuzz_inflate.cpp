// Sanitizer fuzz of csrc/inflate_fast.h against zlib (valid streams of random data / level / strategy, then bit
// flips, truncations and overwrites; exact-size heap buffers so that ASAN sees any access outside the contract):
//   g++ -O1 -g -fsanitize=address,undefined tools/fuzz_inflate.cpp -lz -o /tmp/fuzz_inflate && /tmp/fuzz_inflate
// Last run: 400 valid streams identical, 80,000 corrupted ones: 65,121 refused, 14,374 decoded to different bytes
// (the caller's CRC-32 check catches those), no sanitizer report.
#include "../tredparse_b200/csrc/inflate_fast.h"
#include <zlib.h>
#include <vector>
#include <random>
#include <cstdio>
#include <cstring>
using namespace std;
static vector<uint8_t> deflate_raw(const vector<uint8_t>& in, int level, int strategy) {
    z_stream zs; memset(&zs, 0, sizeof(zs));
    deflateInit2(&zs, level, Z_DEFLATED, -15, 9, strategy);
    vector<uint8_t> out(deflateBound(&zs, in.size()) + 64);
    zs.next_in = (Bytef*)in.data(); zs.avail_in = in.size(); zs.next_out = out.data(); zs.avail_out = out.size();
    deflate(&zs, Z_FINISH); out.resize(zs.total_out); deflateEnd(&zs); return out;
}
int main() {
    mt19937_64 rng(42);
    tredsw_inflate::FastInflater dec;
    long ok = 0, refused = 0, wrong = 0, total = 0;
    for (int round = 0; round < 400; ++round) {
        size_t n = rng() % 70000;
        vector<uint8_t> data(n);
        int mode = round % 5;
        for (size_t i = 0; i < n; ++i) {
            if (mode == 0) data[i] = "ACGT"[rng() & 3];
            else if (mode == 1) data[i] = (uint8_t)(rng() & 255);
            else if (mode == 2) data[i] = "CAG"[i % 3];
            else if (mode == 3) data[i] = (uint8_t)(i < 8 ? rng() : data[i - 1 - (rng() % 8)]);
            else data[i] = (uint8_t)((rng() % 100 < 90) ? 'I' : (33 + rng() % 40));
        }
        int level = (int)(rng() % 10), strategy = (int)(rng() % 5);
        vector<uint8_t> comp = deflate_raw(data, level, strategy);
        // exact-size output buffer + SLACK so ASAN sees any overrun beyond the contract
        vector<uint8_t> out(n + tredsw_inflate::FastInflater::SLACK);
        bool r = dec.inflate(comp.data(), comp.size(), out.data(), n);
        if (!r || memcmp(out.data(), data.data(), n)) { printf("MISMATCH on valid stream round %d n=%zu level=%d strat=%d\n", round, n, level, strategy); return 1; }
        ++ok;
        for (int k = 0; k < 200; ++k) {
            vector<uint8_t> bad = comp;                       // heap copy of exact size: ASAN catches over-reads
            int kind = (int)(rng() % 4);
            if (bad.empty()) break;
            if (kind == 0) bad[rng() % bad.size()] ^= (uint8_t)(1u << (rng() % 8));
            else if (kind == 1) bad.resize(rng() % bad.size());
            else if (kind == 2) for (int j = 0; j < 8 && !bad.empty(); ++j) bad[rng() % bad.size()] = (uint8_t)rng();
            else { size_t a = rng() % bad.size(); for (size_t j = a; j < bad.size() && j < a + 16; ++j) bad[j] = (uint8_t)rng(); }
            size_t want = (rng() % 3 == 0) ? rng() % 80000 : n;
            vector<uint8_t> o2(want + tredsw_inflate::FastInflater::SLACK);
            bool rr = dec.inflate(bad.data(), bad.size(), o2.data(), want);
            ++total;
            if (!rr) ++refused; else if (want != n || memcmp(o2.data(), data.data(), n)) ++wrong;
        }
    }
    printf("valid ok %ld; corrupted: %ld total, %ld refused, %ld decoded-but-different (CRC's job)\n", ok, total, refused, wrong);
    return 0;
}
