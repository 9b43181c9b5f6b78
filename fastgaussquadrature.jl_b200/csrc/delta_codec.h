// delta_codec.h — lossless transport coding of a rule on its way over PCIe (host-returning calls).
//
// Consecutive nodes and weights of a large rule are samples of smooth functions whose step is ~2^24 ulp and
// whose curvature is a fraction of an ulp per step: the SECOND difference of the IEEE bit patterns is a few
// ulp of rounding noise.  A block of 1024 values is sent as four interleaved sequences (stride 4, so that the
// decoder's two running sums live in the four 64-bit lanes of one AVX2 register):
//     [u_0 .. u_3][u_4-u_0 .. u_7-u_3][d2_8 ... d2_{len-1}]     d2_i = (u_i - u_{i-4}) - (u_{i-4} - u_{i-8})
// with u_i the bit pattern (int64) and the d2 stored in the narrowest of 8 / 16 / 32 / 64 bits that holds all
// of them (two's-complement wrap-around arithmetic throughout, so ANY 64-bit content round-trips exactly;
// values that are not smooth — a binade crossing, the first few million weights, the centre node — simply take a
// wider class).  Typical cost at n = 1e9: 1.03 bytes per value instead of 8.  The device packs (codec.cu), host
// threads unpack straight into the caller's arrays (capi.cu); nothing of the quadrature algorithm runs on the
// host — this is a byte codec, and tests check decode(encode(v)) == v bit for bit.
#ifndef FGQ_DELTA_CODEC_H_
#define FGQ_DELTA_CODEC_H_

#include <stdint.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#if defined(__CUDACC__)
#define FGQ_CODEC_HD __host__ __device__ __forceinline__
#else
#define FGQ_CODEC_HD static inline
#endif

namespace fgq {

constexpr int CODEC_BLOCK = 1024;        // values per block
constexpr int CODEC_STRIDE = 4;          // interleaved sequences per block
constexpr int CODEC_HEADER8 = 8;         // u_0..u_3 and u_4-u_0..u_7-u_3, in 8-byte units

// width class of one second difference: 0 -> int8, 1 -> int16, 2 -> int32, 3 -> int64
FGQ_CODEC_HD int codec_class_of(int64_t d2) {
    if (d2 >= -128 && d2 <= 127) return 0;
    if (d2 >= -32768 && d2 <= 32767) return 1;
    if (d2 >= -2147483648LL && d2 <= 2147483647LL) return 2;
    return 3;
}

// payload size of a block of `len` values in class `cls`, in 8-byte units (header included)
FGQ_CODEC_HD uint32_t codec_block_size8(int len, int cls) {
    const int nd2 = len > 2 * CODEC_STRIDE ? len - 2 * CODEC_STRIDE : 0;
    return (uint32_t)(CODEC_HEADER8 + (((int64_t)nd2 << cls) + 7) / 8);
}

// directory entry of a block: (offset in 8-byte units << 2) | class
FGQ_CODEC_HD uint32_t codec_dir_entry(uint32_t off8, int cls) { return (off8 << 2) | (uint32_t)cls; }

// worst-case stream size of `cnt` values, bytes
static inline int64_t codec_worst_bytes(int64_t cnt) {
    const int64_t nb = (cnt + CODEC_BLOCK - 1) / CODEC_BLOCK;
    return cnt * 8 + nb * 8 * CODEC_HEADER8;
}

// ---- host side ------------------------------------------------------------------------------------------

// second difference i of a block (stride 4), from its bit patterns; i >= 8
FGQ_CODEC_HD uint64_t codec_d2(const uint64_t* v, int i) {
    return (v[i] - v[i - CODEC_STRIDE]) - (v[i - CODEC_STRIDE] - v[i - 2 * CODEC_STRIDE]);
}

static inline uint64_t codec_load_d2(const uint8_t* q, int cls, int k) {  // k-th stored second difference, sign-extended
    switch (cls) {
        case 0: return (uint64_t)(int64_t)(int8_t)q[k];
        case 1: { int16_t t; memcpy(&t, q + 2 * (size_t)k, 2); return (uint64_t)(int64_t)t; }
        case 2: { int32_t t; memcpy(&t, q + 4 * (size_t)k, 4); return (uint64_t)(int64_t)t; }
        default: { uint64_t t; memcpy(&t, q + 8 * (size_t)k, 8); return t; }
    }
}

// Portable decoder: one block (len values) from `p` (class cls) into out[0..len).
static inline void codec_decode_block_scalar(const uint8_t* p, int cls, int len, int64_t* out) {
    uint64_t v[CODEC_STRIDE], d1[CODEC_STRIDE];
    memcpy(v, p, 32);
    memcpy(d1, p + 32, 32);
    for (int i = 0; i < len && i < CODEC_STRIDE; ++i) out[i] = (int64_t)v[i];
    for (int s = 0; s < CODEC_STRIDE; ++s) v[s] += d1[s];
    for (int i = CODEC_STRIDE; i < len && i < 2 * CODEC_STRIDE; ++i) out[i] = (int64_t)v[i - CODEC_STRIDE];
    const uint8_t* q = p + 64;
    for (int i = 2 * CODEC_STRIDE; i < len; ++i) {
        const int s = i & (CODEC_STRIDE - 1);
        d1[s] += codec_load_d2(q, cls, i - 2 * CODEC_STRIDE);
        v[s] += d1[s];
        out[i] = (int64_t)v[s];
    }
}

#if defined(__x86_64__)
// AVX2 decoder: the four interleaved sequences are the four lanes; two vector adds and one store per 4 values.
__attribute__((target("avx2"))) static inline void codec_decode_block_avx2(const uint8_t* p, int cls, int len, int64_t* out) {
    if (len < 2 * CODEC_STRIDE) {
        codec_decode_block_scalar(p, cls, len, out);
        return;
    }
    __m256i v = _mm256_loadu_si256((const __m256i*)p);
    __m256i d1 = _mm256_loadu_si256((const __m256i*)(p + 32));
    _mm256_storeu_si256((__m256i*)out, v);
    v = _mm256_add_epi64(v, d1);
    _mm256_storeu_si256((__m256i*)(out + 4), v);
    const uint8_t* q = p + 64;
    const int full = len & ~3;  // whole groups of 4
    switch (cls) {
        case 0:
            for (int i = 8; i < full; i += 4) {
                int32_t four;
                memcpy(&four, q + (i - 8), 4);
                d1 = _mm256_add_epi64(d1, _mm256_cvtepi8_epi64(_mm_cvtsi32_si128(four)));
                v = _mm256_add_epi64(v, d1);
                _mm256_storeu_si256((__m256i*)(out + i), v);
            }
            break;
        case 1:
            for (int i = 8; i < full; i += 4) {
                d1 = _mm256_add_epi64(d1, _mm256_cvtepi16_epi64(_mm_loadl_epi64((const __m128i*)(q + 2 * (size_t)(i - 8)))));
                v = _mm256_add_epi64(v, d1);
                _mm256_storeu_si256((__m256i*)(out + i), v);
            }
            break;
        case 2:
            for (int i = 8; i < full; i += 4) {
                d1 = _mm256_add_epi64(d1, _mm256_cvtepi32_epi64(_mm_loadu_si128((const __m128i*)(q + 4 * (size_t)(i - 8)))));
                v = _mm256_add_epi64(v, d1);
                _mm256_storeu_si256((__m256i*)(out + i), v);
            }
            break;
        default:
            for (int i = 8; i < full; i += 4) {
                d1 = _mm256_add_epi64(d1, _mm256_loadu_si256((const __m256i*)(q + 8 * (size_t)(i - 8))));
                v = _mm256_add_epi64(v, d1);
                _mm256_storeu_si256((__m256i*)(out + i), v);
            }
    }
    if (full < len) {  // the last, partial group (only the final block of a rule): scalar lanes
        uint64_t vv[4], dd[4];
        _mm256_storeu_si256((__m256i*)vv, v);
        _mm256_storeu_si256((__m256i*)dd, d1);
        for (int i = full; i < len; ++i) {
            const int s = i & 3;
            dd[s] += codec_load_d2(q, cls, i - 8);
            vv[s] += dd[s];
            out[i] = (int64_t)vv[s];
        }
    }
}
#endif

static inline void codec_decode_block(const uint8_t* p, int cls, int len, int64_t* out) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) {
        codec_decode_block_avx2(p, cls, len, out);
        return;
    }
#endif
    codec_decode_block_scalar(p, cls, len, out);
}

// Host reference encoder (tests only: the product packs on the device).  payload must hold
// codec_worst_bytes(cnt) bytes, dir ceil(cnt / 1024) entries; returns the stream size in bytes.
static inline int64_t codec_encode_host(const int64_t* u, int64_t cnt, uint8_t* payload, uint32_t* dir) {
    const int64_t nb = (cnt + CODEC_BLOCK - 1) / CODEC_BLOCK;
    uint32_t off8 = 0;
    for (int64_t b = 0; b < nb; ++b) {
        const int64_t base = b * CODEC_BLOCK;
        const int len = (int)(cnt - base < CODEC_BLOCK ? cnt - base : CODEC_BLOCK);
        const uint64_t* v = (const uint64_t*)(u + base);
        int cls = 0;
        for (int i = 2 * CODEC_STRIDE; i < len; ++i) {
            const int c = codec_class_of((int64_t)codec_d2(v, i));
            cls = c > cls ? c : cls;
        }
        dir[b] = codec_dir_entry(off8, cls);
        uint8_t* p = payload + (size_t)off8 * 8;
        const uint32_t size8 = codec_block_size8(len, cls);
        memset(p, 0, (size_t)size8 * 8);
        for (int i = 0; i < CODEC_STRIDE; ++i) {
            const uint64_t a = i < len ? v[i] : 0;
            const uint64_t d = i + CODEC_STRIDE < len ? v[i + CODEC_STRIDE] - v[i] : 0;
            memcpy(p + 8 * i, &a, 8);
            memcpy(p + 32 + 8 * i, &d, 8);
        }
        for (int i = 2 * CODEC_STRIDE; i < len; ++i) {
            const uint64_t d2 = codec_d2(v, i);
            memcpy(p + 64 + ((size_t)(i - 2 * CODEC_STRIDE) << cls), &d2, (size_t)1 << cls);  // little-endian truncation
        }
        off8 += size8;
    }
    return (int64_t)off8 * 8;
}
}  // namespace fgq
#endif
