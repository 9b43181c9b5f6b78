// delta_codec.h — lossless transport coding of a rule on its way over PCIe (host-returning calls).
//
// Consecutive nodes and weights of a large rule are samples of smooth functions whose step is ~2^24 ulp and
// whose curvature is a fraction of an ulp per step: the SECOND difference of the IEEE bit patterns is a few
// ulp of rounding noise.  A block of 1024 values is therefore sent as
//     [u_0][d1_1][d2_2 ... d2_{len-1}]      u_i = bit pattern (int64), d1_i = u_i - u_{i-1}, d2_i = d1_i - d1_{i-1}
// with the d2 stored in the narrowest of 8 / 16 / 32 / 64 bits that holds all of them (two's-complement
// wrap-around arithmetic throughout, so ANY 64-bit content round-trips exactly; values that are not smooth —
// a binade crossing, the first few thousand weights, the centre node — simply take a wider class).  Typical
// cost at n = 1e9: 1.0 byte per value instead of 8.  The device packs (codec.cu), host threads unpack straight
// into the caller's arrays (capi.cu); nothing of the quadrature algorithm runs on the host — this is a byte
// codec, and tests check decode(encode(v)) == v bit for bit.
#ifndef FGQ_DELTA_CODEC_H_
#define FGQ_DELTA_CODEC_H_

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FGQ_CODEC_HD __host__ __device__ __forceinline__
#else
#define FGQ_CODEC_HD static inline
#endif

namespace fgq {

constexpr int CODEC_BLOCK = 1024;        // values per block
constexpr int CODEC_HEADER8 = 2;         // u_0 and d1_1, in 8-byte units

// width class of one second difference: 0 -> int8, 1 -> int16, 2 -> int32, 3 -> int64
FGQ_CODEC_HD int codec_class_of(int64_t d2) {
    if (d2 >= -128 && d2 <= 127) return 0;
    if (d2 >= -32768 && d2 <= 32767) return 1;
    if (d2 >= -2147483648LL && d2 <= 2147483647LL) return 2;
    return 3;
}

// payload size of a block of `len` values in class `cls`, in 8-byte units (header included)
FGQ_CODEC_HD uint32_t codec_block_size8(int len, int cls) {
    const int nd2 = len > 2 ? len - 2 : 0;
    return (uint32_t)(CODEC_HEADER8 + (((int64_t)nd2 << cls) + 7) / 8);
}

// directory entry of a block: (offset in 8-byte units << 2) | class
FGQ_CODEC_HD uint32_t codec_dir_entry(uint32_t off8, int cls) { return (off8 << 2) | (uint32_t)cls; }

// worst-case stream size of `cnt` values, bytes
static inline int64_t codec_worst_bytes(int64_t cnt) {
    const int64_t nb = (cnt + CODEC_BLOCK - 1) / CODEC_BLOCK;
    return cnt * 8 + nb * 16;
}

// ---- host side ------------------------------------------------------------------------------------------

// Decodes one block (len values) from `p` (class cls) into out[0..len).
static inline void codec_decode_block(const uint8_t* p, int cls, int len, int64_t* out) {
    uint64_t v, d1;
    memcpy(&v, p, 8);
    memcpy(&d1, p + 8, 8);
    out[0] = (int64_t)v;
    if (len == 1) return;
    v += d1;
    out[1] = (int64_t)v;
    const uint8_t* q = p + 16;
    switch (cls) {
        case 0: {
            const int8_t* s = (const int8_t*)q;
            for (int i = 2; i < len; ++i) {
                d1 += (uint64_t)(int64_t)s[i - 2];
                v += d1;
                out[i] = (int64_t)v;
            }
            break;
        }
        case 1: {
            for (int i = 2; i < len; ++i) {
                int16_t t;
                memcpy(&t, q + 2 * (i - 2), 2);
                d1 += (uint64_t)(int64_t)t;
                v += d1;
                out[i] = (int64_t)v;
            }
            break;
        }
        case 2: {
            for (int i = 2; i < len; ++i) {
                int32_t t;
                memcpy(&t, q + 4 * (i - 2), 4);
                d1 += (uint64_t)(int64_t)t;
                v += d1;
                out[i] = (int64_t)v;
            }
            break;
        }
        default: {
            for (int i = 2; i < len; ++i) {
                uint64_t t;
                memcpy(&t, q + 8 * (size_t)(i - 2), 8);
                d1 += t;
                v += d1;
                out[i] = (int64_t)v;
            }
        }
    }
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
        for (int i = 2; i < len; ++i) {
            const int64_t d2 = (int64_t)((v[i] - v[i - 1]) - (v[i - 1] - v[i - 2]));
            const int c = codec_class_of(d2);
            cls = c > cls ? c : cls;
        }
        dir[b] = codec_dir_entry(off8, cls);
        uint8_t* p = payload + (size_t)off8 * 8;
        const uint64_t d1 = len > 1 ? v[1] - v[0] : 0;
        memcpy(p, &v[0], 8);
        memcpy(p + 8, &d1, 8);
        const uint32_t size8 = codec_block_size8(len, cls);
        if (size8 > CODEC_HEADER8) memset(p + (size_t)(size8 - 1) * 8, 0, 8);  // padding bytes of the last word
        for (int i = 2; i < len; ++i) {
            const uint64_t d2 = (v[i] - v[i - 1]) - (v[i - 1] - v[i - 2]);
            memcpy(p + 16 + ((size_t)(i - 2) << cls), &d2, (size_t)1 << cls);  // little-endian truncation
        }
        off8 += size8;
    }
    return (int64_t)off8 * 8;
}

}  // namespace fgq
#endif
