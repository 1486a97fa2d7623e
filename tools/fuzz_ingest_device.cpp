// Sanitizer fuzz of the GPU ingest's device code (csrc/ingest_device.cuh, compiled for the host: NL = 1) under
// corruption — the contract the kernels rely on for memory safety:
//   * inflate_block reads at most 16 bytes behind the compressed input and writes only out[0, out_len);
//   * walk_fetch / select_record / emit_read / the pairing helpers stay inside [0, run_end) (+ the 8 bytes of slack the
//     buffers are allocated with) whatever the bytes say.
// Exact-size heap buffers, so that ASan sees any access outside the contract.
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined tools/fuzz_ingest_device.cpp -lz -o /tmp/fuzz_ingest && /tmp/fuzz_ingest tests/golden/t001.mini.bam
// Last run (tests/golden/t001.mini.bam): 29 blocks / 6,381 records; 58,000 corrupted DEFLATE streams (52,181 refused,
// 5,814 decoded to other bytes = what the CRC-32 check catches), 3,000 corrupted record streams = 18.4 M records walked,
// selected, hashed and emitted (125 walks flagged): no sanitizer report.  (The first run found the bit reader of a
// corrupt stream 12 bytes behind its input — 4 more than documented; the buffers had the room, the contract now says 16.)
#include "../tredparse_b200/csrc/ingest_device.cuh"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
using namespace tredsw_gi;
using namespace std;

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: fuzz_ingest file.bam\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    vector<uint8_t> file(n);
    if (fread(file.data(), 1, n, f) != (size_t)n) return 2;
    fclose(f);
    mt19937_64 rng(7);
    vector<uint16_t> tabs(TAB_ENTRIES);
    uint8_t lens[320], sub_need[1 << LIT_ROOT];
    uint16_t sub_base[1 << LIT_ROOT];
    vector<uint8_t> stream;                       // the inflated record stream of the whole file
    long refused = 0, wrong = 0, blocks = 0;
    for (long o = 0; o + 18 <= n;) {
        const int xlen = file[o + 10] | (file[o + 11] << 8), bsize = (file[o + 16] | (file[o + 17] << 8)) + 1;
        const int clen = bsize - xlen - 19;
        uint32_t isize; memcpy(&isize, &file[o + bsize - 4], 4);
        if (isize) {
            // exact-size input (+16: the documented read-ahead of the bit reader) and output
            vector<uint8_t> in(file.begin() + o + 12 + xlen, file.begin() + o + 12 + xlen + clen);
            in.resize(clen + 16, 0);
            vector<uint8_t> out(isize);
            if (!inflate_block<1>(in.data(), clen, out.data(), isize, TabRef{tabs.data(), 1, true}, lens, sub_need, sub_base, 0)) { printf("valid block refused at %ld\n", o); return 1; }
            stream.insert(stream.end(), out.begin(), out.end());
            for (int k = 0; k < 2000; ++k) {      // corrupted copies of this block
                vector<uint8_t> bad(in);
                const int how = (int)(rng() % 3);
                int blen = clen;
                if (how == 0) for (int j = 0; j < 1 + (int)(rng() % 3); ++j) bad[rng() % clen] ^= (uint8_t)(1u << (rng() % 8));
                else if (how == 1) { blen = 1 + (int)(rng() % clen); bad.resize(blen + 16); memset(bad.data() + blen, 0, 16); }
                else for (int j = 0; j < 8; ++j) bad[rng() % clen] = (uint8_t)rng();
                vector<uint8_t> o2(isize);
                const bool r = inflate_block<1>(bad.data(), blen, o2.data(), isize, TabRef{tabs.data(), 1, true}, lens, sub_need, sub_base, 0);
                if (!r) ++refused; else if (memcmp(o2.data(), out.data(), isize)) ++wrong;
            }
            ++blocks;
        }
        o += bsize;
    }
    // records: skip the header, then corrupt the stream and walk / select / pair / emit over it
    size_t p = 4; int32_t l_text; memcpy(&l_text, &stream[p], 4); p += 4 + l_text;
    int32_t n_ref; memcpy(&n_ref, &stream[p], 4); p += 4;
    for (int i = 0; i < n_ref; ++i) { int32_t l; memcpy(&l, &stream[p], 4); p += 4 + l + 4; }
    const size_t first = p;
    long records = 0, walked = 0, flagged = 0;
    for (int round = 0; round < 3000; ++round) {
        // exact-size copy (+8 slack as allocated by the pipeline)
        vector<uint8_t> buf(stream.size() + 8, 0);
        memcpy(buf.data(), stream.data(), stream.size());
        if (round) for (int j = 0; j < 1 + (int)(rng() % 6); ++j) {
            const size_t at = first + rng() % (stream.size() - first);
            if (rng() & 1) buf[at] ^= (uint8_t)(1u << (rng() % 8)); else { uint32_t v = (uint32_t)rng(); memcpy(&buf[at], &v, min<size_t>(4, stream.size() - at)); }
        }
        Chunk ch{(int64_t)first, (int64_t)stream.size(), (int64_t)stream.size()};
        Fetch fe{0, round % 2, 0, 0, 1, 0, 0, 1ll << 40};
        int32_t tid0; memcpy(&tid0, &buf[first + 4], 4);
        fe.tid = tid0;
        ProblemParams q{tid0, 1000, 0, 1ll << 40, 0, 1ll << 40, 0, 1ll << 40, 100, 200};
        ProblemCounts cnt{0, 0, 0, 0};
        vector<int64_t> pos;
        DirectReader rd{buf.data()};
        const bool ok = walk_fetch(rd, fe, &ch, [&](int64_t at) { pos.push_back(at); });
        if (!ok) ++flagged;
        vector<int8_t> rb; vector<char> nm;
        for (int64_t at : pos) {
            Mate m{};
            const RecOut o = select_record(buf.data(), at, fe, q, &cnt, &m);
            if (o.emit) {
                rb.assign(o.bases, 0); nm.assign(o.name_bytes, 0);
                emit_read(buf.data(), at, rb.data(), nm.data(), 0, 1);
            }
            if (o.pe) (void)name_hash(buf.data(), at);
            ++walked;
        }
        if (pos.size() > 1) (void)same_name(buf.data(), pos[0], pos[pos.size() / 2]);
        if (!round) records = (long)pos.size();
    }
    printf("%ld blocks, %ld records; corrupted streams: %ld refused, %ld decoded to other bytes; %ld records walked over corrupted streams, %ld walks flagged\n",
           blocks, records, refused, wrong, walked, flagged);
    return 0;
}
