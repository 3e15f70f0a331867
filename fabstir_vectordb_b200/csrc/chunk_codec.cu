// chunk_codec.cu — VectorChunk CBOR <-> dense (ids, row-major fp32) arrays.  Host code only
// (compiled with the rest of the library; no CUDA call).  Contract: include/fvdb_chunk.h.
//
// Follows RFC 8949 for the data model and serde_cbor 0.11's conventions for `VectorChunk`
// (src/core/chunk.rs:37-43, to_cbor/from_cbor :78-86): struct = map with text keys in field order,
// usize = unsigned int in its shortest form, [u8; 32] = array of 32 unsigned ints, Vec<f32> = array
// of floats where a value that survives the round trip through binary16 is written as 0xf9.
// (serde_cbor is an un-vendored dependency, Cargo.toml:17: its behaviour is restated, and the reader
// accepts every well-formed spelling so that reading never depends on that restatement.)
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fvdb.h"
#include "../../include/fvdb_chunk.h"

#define FVDB_EXPORT extern "C"   /* default visibility comes from the header's pragma */

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

// ---- binary16 <-> binary32 ---------------------------------------------------------------------
uint32_t f32_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
float bits_f32(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// round-to-nearest-even, overflow to infinity (what half::f16::from_f32 does)
uint16_t f32_to_f16(float f) {
    const uint32_t x = f32_bits(f);
    const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    const uint32_t exp = (x >> 23) & 0xffu;
    uint32_t man = x & 0x7fffffu;
    if (exp == 0xffu) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0u));
    const int e = (int)exp - 127 + 15;
    if (e >= 31) return (uint16_t)(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return sign;
        man |= 0x800000u;
        const int shift = 14 - e;
        uint32_t h = man >> shift;
        const uint32_t rem = man & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1u))) ++h;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((uint32_t)e << 10) | (man >> 13);
    const uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
    return (uint16_t)(sign | h);
}

float f16_to_f32(uint16_t h) {
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    const uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    if (exp == 0) {
        if (man == 0) return bits_f32(sign);
        int e = -1;
        do { man <<= 1; ++e; } while (!(man & 0x400u));
        return bits_f32(sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13));
    }
    if (exp == 31) return bits_f32(sign | 0x7f800000u | (man << 13));
    return bits_f32(sign | ((exp + 127 - 15) << 23) | (man << 13));
}

// ---- reader --------------------------------------------------------------------------------------
struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    const char* why = "";

    bool bad(const char* w) { if (ok) { ok = false; why = w; } return false; }
    size_t left() const { return (size_t)(end - p); }

    // item head: major type, additional info, argument.  Tags (major 6) are consumed here.
    bool head(unsigned& major, unsigned& ai, uint64_t& arg) {
        for (;;) {
            if (left() < 1) return bad("truncated item head");
            const uint8_t b = *p++;
            major = b >> 5;
            ai = b & 31u;
            arg = ai;
            if (ai >= 24 && ai <= 27) {
                const size_t n = (size_t)1 << (ai - 24);
                if (left() < n) return bad("truncated item argument");
                arg = 0;
                for (size_t i = 0; i < n; ++i) arg = (arg << 8) | *p++;
            } else if (ai >= 28 && ai <= 30) {
                return bad("reserved additional information");
            }
            if (major != 6) return true;
            if (ai == 31) return bad("indefinite tag");
        }
    }
    bool peek_break() const { return left() >= 1 && *p == 0xffu; }

    bool skip(int depth = 0) {
        if (depth > 64) return bad("nesting too deep");
        unsigned major, ai;
        uint64_t arg;
        if (!head(major, ai, arg)) return false;
        switch (major) {
        case 0: case 1: return ai != 31 || bad("indefinite integer");
        case 2: case 3:
            if (ai == 31) {
                while (!peek_break()) {
                    unsigned m2, a2;
                    uint64_t n2;
                    if (!head(m2, a2, n2)) return false;
                    if (m2 != major || a2 == 31) return bad("bad chunk of an indefinite string");
                    if (left() < n2) return bad("truncated string");
                    p += n2;
                }
                ++p;
                return true;
            }
            if (left() < arg) return bad("truncated string");
            p += arg;
            return true;
        case 4: case 5: {
            const uint64_t per = major == 5 ? 2 : 1;
            if (ai == 31) {
                while (!peek_break()) {
                    for (uint64_t j = 0; j < per; ++j) if (!skip(depth + 1)) return false;
                    if (left() < 1) return bad("truncated container");
                }
                ++p;
                return true;
            }
            if (arg > left()) return bad("container longer than the input");
            for (uint64_t i = 0; i < arg * per; ++i) if (!skip(depth + 1)) return false;
            return true;
        }
        default:  // 7: simple values and floats (the argument bytes are already consumed)
            return ai != 31 || bad("unexpected break");
        }
    }

    bool uint(uint64_t& v) {
        unsigned major, ai;
        if (!head(major, ai, v)) return false;
        return (major == 0 && ai != 31) || bad("expected an unsigned integer");
    }

    // text string into out (definite or chunked)
    bool text(std::string& out) {
        unsigned major, ai;
        uint64_t arg;
        if (!head(major, ai, arg)) return false;
        if (major != 3) return bad("expected a text string");
        out.clear();
        if (ai == 31) {
            while (!peek_break()) {
                unsigned m2, a2;
                uint64_t n2;
                if (!head(m2, a2, n2)) return false;
                if (m2 != 3 || a2 == 31) return bad("bad chunk of an indefinite text string");
                if (left() < n2) return bad("truncated text string");
                out.append(reinterpret_cast<const char*>(p), (size_t)n2);
                p += n2;
            }
            ++p;
            return true;
        }
        if (left() < arg) return bad("truncated text string");
        out.assign(reinterpret_cast<const char*>(p), (size_t)arg);
        p += arg;
        return true;
    }

    // one number as f32: half / single / double float, or an integer
    bool number(float& f) {
        if (left() >= 5 && *p == 0xfau) {  // the common case
            f = bits_f32(((uint32_t)p[1] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 8) | p[4]);
            p += 5;
            return true;
        }
        unsigned major, ai;
        uint64_t arg;
        if (!head(major, ai, arg)) return false;
        if (major == 7 && ai == 25) { f = f16_to_f32((uint16_t)arg); return true; }
        if (major == 7 && ai == 26) { f = bits_f32((uint32_t)arg); return true; }
        if (major == 7 && ai == 27) { double d; std::memcpy(&d, &arg, 8); f = (float)d; return true; }
        if (major == 0 && ai != 31) { f = (float)arg; return true; }
        if (major == 1 && ai != 31) { f = -1.0f - (float)arg; return true; }
        return bad("expected a number");
    }

    // container head: returns the length, or UINT64_MAX for an indefinite one
    bool container(unsigned want_major, uint64_t& n, const char* what) {
        unsigned major, ai;
        if (!head(major, ai, n)) return false;
        if (major != want_major) return bad(what);
        if (ai == 31) n = UINT64_MAX;
        else if (n > left()) return bad("container longer than the input");
        return true;
    }
    // true while the container has another element (consumes the break of an indefinite one)
    bool more(uint64_t n, uint64_t i) {
        if (n != UINT64_MAX) return i < n;
        if (left() < 1) { bad("truncated container"); return false; }
        if (*p == 0xffu) { ++p; return false; }
        return true;
    }
};

// ---- writer --------------------------------------------------------------------------------------
struct Writer {
    uint8_t* out;    // nullptr: size only
    size_t cap;
    size_t n = 0;
    bool overflow = false;

    void byte(uint8_t b) {
        if (out) { if (n < cap) out[n] = b; else overflow = true; }
        ++n;
    }
    void head(unsigned major, uint64_t arg) {
        const uint8_t m = (uint8_t)(major << 5);
        if (arg < 24) byte((uint8_t)(m | arg));
        else if (arg <= 0xffu) { byte(m | 24); byte((uint8_t)arg); }
        else if (arg <= 0xffffu) { byte(m | 25); byte((uint8_t)(arg >> 8)); byte((uint8_t)arg); }
        else if (arg <= 0xffffffffull) { byte(m | 26); for (int s = 24; s >= 0; s -= 8) byte((uint8_t)(arg >> s)); }
        else { byte(m | 27); for (int s = 56; s >= 0; s -= 8) byte((uint8_t)(arg >> s)); }
    }
    void text(const char* s, size_t len) {
        head(3, len);
        for (size_t i = 0; i < len; ++i) byte((uint8_t)s[i]);
    }
    // serde_cbor 0.11 Serializer::serialize_f32: infinities and NaN as fixed half floats, a value that
    // survives f32 -> f16 -> f32 as a half float, everything else as a single float
    void f32(float v) {
        const uint32_t b = f32_bits(v);
        if ((b & 0x7fffffffu) == 0x7f800000u) { byte(0xf9); byte((b >> 31) ? 0xfc : 0x7c); byte(0x00); return; }
        if ((b & 0x7fffffffu) > 0x7f800000u) { byte(0xf9); byte(0x7e); byte(0x00); return; }
        const uint16_t h = f32_to_f16(v);
        if (f16_to_f32(h) == v) { byte(0xf9); byte((uint8_t)(h >> 8)); byte((uint8_t)h); return; }
        byte(0xfa);
        for (int s = 24; s >= 0; s -= 8) byte((uint8_t)(b >> s));
    }
};

}  // namespace

FVDB_EXPORT const char* fvdb_chunk_last_error(void) { return g_err.c_str(); }

FVDB_EXPORT int fvdb_chunk_decode(const uint8_t* cbor, size_t len, fvdb_chunk_info* info, uint8_t* out_ids,
                                  float* out_rows, uint64_t cap_vectors) {
    g_err.clear();
    if (!cbor || !info) return fail(FVDB_ERR_INVALID_ARG, "fvdb_chunk_decode: null argument");
    if ((out_ids == nullptr) != (out_rows == nullptr))
        return fail(FVDB_ERR_INVALID_ARG, "fvdb_chunk_decode: give both output arrays or neither");
    std::memset(info, 0, sizeof(*info));
    Reader r{cbor, cbor + len};
    uint64_t n_fields;
    if (!r.container(5, n_fields, "a VectorChunk is a map")) return fail(FVDB_ERR_CHUNK_LOAD, std::string("chunk: ") + r.why);
    bool have_id = false, have_start = false, have_end = false, have_vectors = false;
    uint64_t count = 0;
    uint32_t dim = 0;
    bool dim_known = false;
    std::string key, tmp_text;
    std::vector<float> tmp_row;
    for (uint64_t f = 0; r.more(n_fields, f); ++f) {
        if (!r.text(key)) break;
        if (key == "chunk_id") {
            if (have_id) { r.bad("duplicate field chunk_id"); break; }
            if (!r.text(tmp_text)) break;
            std::snprintf(info->chunk_id, sizeof(info->chunk_id), "%s", tmp_text.c_str());
            have_id = true;
        } else if (key == "start_idx") {
            if (have_start) { r.bad("duplicate field start_idx"); break; }
            if (!r.uint(info->start_idx)) break;
            have_start = true;
        } else if (key == "end_idx") {
            if (have_end) { r.bad("duplicate field end_idx"); break; }
            if (!r.uint(info->end_idx)) break;
            have_end = true;
        } else if (key == "vectors") {
            if (have_vectors) { r.bad("duplicate field vectors"); break; }
            have_vectors = true;
            uint64_t n_vec;
            if (!r.container(5, n_vec, "`vectors` is a map")) break;
            for (uint64_t i = 0; r.more(n_vec, i); ++i) {
                if (out_ids && count >= cap_vectors)
                    return fail(FVDB_ERR_INVALID_ARG, "fvdb_chunk_decode: more vectors than cap_vectors");
                // ---- key: the VectorId, array(32) of u8 (or a 32-byte byte string)
                uint8_t id[32];
                {
                    unsigned major, ai;
                    uint64_t arg;
                    if (!r.head(major, ai, arg)) break;
                    if (major == 2 && ai != 31 && arg == 32) {
                        if (r.left() < 32) { r.bad("truncated id"); break; }
                        std::memcpy(id, r.p, 32);
                        r.p += 32;
                    } else if (major == 4) {
                        const uint64_t n_id = ai == 31 ? UINT64_MAX : arg;
                        uint64_t j = 0;
                        for (; r.more(n_id, j); ++j) {
                            uint64_t b;
                            if (!r.uint(b)) break;
                            if (b > 255 || j >= 32) { r.bad("a VectorId is 32 bytes"); break; }
                            id[j] = (uint8_t)b;
                        }
                        if (!r.ok) break;
                        if (j != 32) { r.bad("a VectorId is 32 bytes"); break; }
                    } else { r.bad("a VectorId is an array of 32 bytes"); break; }
                }
                // ---- value: the vector
                uint64_t n_el;
                if (!r.container(4, n_el, "a vector is an array of floats")) break;
                float* dst = nullptr;
                if (n_el != UINT64_MAX) {
                    if (!dim_known) {
                        if (n_el > 0xffffffffull) { r.bad("vector too long"); break; }
                        dim = (uint32_t)n_el;
                        dim_known = true;
                    } else if (n_el != dim) {
                        return fail(FVDB_ERR_INCONSISTENT_DIM, "chunk: vectors of different lengths");
                    }
                    if (out_rows) dst = out_rows + (size_t)count * dim;
                    float scratch;
                    for (uint64_t j = 0; j < n_el; ++j)
                        if (!r.number(dst ? dst[j] : scratch)) break;
                    if (!r.ok) break;
                } else {
                    tmp_row.clear();
                    for (uint64_t j = 0; r.more(n_el, j); ++j) {
                        float v;
                        if (!r.number(v)) break;
                        tmp_row.push_back(v);
                    }
                    if (!r.ok) break;
                    if (!dim_known) { dim = (uint32_t)tmp_row.size(); dim_known = true; }
                    else if (tmp_row.size() != dim) return fail(FVDB_ERR_INCONSISTENT_DIM, "chunk: vectors of different lengths");
                    if (out_rows && dim) std::memcpy(out_rows + (size_t)count * dim, tmp_row.data(), (size_t)dim * sizeof(float));
                }
                if (out_ids) std::memcpy(out_ids + (size_t)count * 32, id, 32);
                ++count;
            }
            if (!r.ok) break;
        } else {
            if (!r.skip()) break;  // a field this version does not know
        }
    }
    if (!r.ok) return fail(FVDB_ERR_CHUNK_LOAD, std::string("chunk: ") + r.why);
    if (!have_id || !have_start || !have_end || !have_vectors)
        return fail(FVDB_ERR_CHUNK_LOAD, "chunk: missing field (chunk_id, start_idx, end_idx, vectors)");
    if (r.left() != 0) return fail(FVDB_ERR_CHUNK_LOAD, "chunk: trailing bytes after the VectorChunk");
    info->n_vectors = count;
    info->dim = dim;
    return FVDB_OK;
}

FVDB_EXPORT int fvdb_chunk_encode(const char* chunk_id, uint64_t start_idx, uint64_t end_idx, const uint8_t* ids,
                                  const float* rows, uint64_t n, uint32_t dim, uint8_t* out, size_t cap,
                                  size_t* out_len) {
    g_err.clear();
    if (!chunk_id || !out_len || (n && (!ids || (dim && !rows))))
        return fail(FVDB_ERR_INVALID_ARG, "fvdb_chunk_encode: null argument");
    Writer w{out, cap};
    w.head(5, 4);
    w.text("chunk_id", 8);
    w.text(chunk_id, std::strlen(chunk_id));
    w.text("start_idx", 9);
    w.head(0, start_idx);
    w.text("end_idx", 7);
    w.head(0, end_idx);
    w.text("vectors", 7);
    w.head(5, n);
    for (uint64_t i = 0; i < n; ++i) {
        w.head(4, 32);
        for (int j = 0; j < 32; ++j) w.head(0, ids[(size_t)i * 32 + j]);
        w.head(4, dim);
        for (uint32_t j = 0; j < dim; ++j) w.f32(rows[(size_t)i * dim + j]);
    }
    *out_len = w.n;
    if (w.overflow) return fail(FVDB_ERR_INVALID_ARG, "fvdb_chunk_encode: output buffer too small");
    return FVDB_OK;
}
