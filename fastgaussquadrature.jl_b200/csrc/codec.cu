// codec.cu — device side of the transport coding of delta_codec.h: packs freshly generated nodes and
// weights (still in HBM) into second-difference streams before they cross PCIe.
//
// Three launches per chunk, both arrays (x, w) at once through gridDim.y:
//   codec_class_kernel    one 256-thread block per 1024 values: width class of the block -> its size
//   codec_offsets_kernel  one block: exclusive scan of the sizes -> directory, total stream size
//   codec_pack_kernel     one block per 1024 values: header + packed second differences at its offset
// Reads 16 B/node of HBM again and writes ~2 B/node: ~3 us per 2^24-node chunk, against the ~100 us the
// uncompressed chunk needs on the PCIe link.
#include "common.cuh"
#include "delta_codec.h"

namespace fgq {

constexpr int CODEC_THREADS = 256;  // 4 values per thread

// the block's values as bit patterns, in shared memory
__device__ __forceinline__ int codec_load_block(const uint64_t* __restrict__ v, int64_t cnt, int64_t b, uint64_t* sh) {
    const int64_t base = b * CODEC_BLOCK;
    const int len = (int)(cnt - base < CODEC_BLOCK ? cnt - base : CODEC_BLOCK);
    for (int i = threadIdx.x; i < len; i += CODEC_THREADS) sh[i] = v[base + i];
    __syncthreads();
    return len;
}

__global__ void __launch_bounds__(CODEC_THREADS)
codec_class_kernel(CodecArrays a, int64_t cnt, int64_t nb) {
    __shared__ uint64_t sh[CODEC_BLOCK];
    __shared__ int s_cls;
    const int arr = blockIdx.y;
    const int64_t b = blockIdx.x;
    if (threadIdx.x == 0) s_cls = 0;
    const int len = codec_load_block((const uint64_t*)a.values[arr], cnt, b, sh);
    int cls = 0;
    for (int i = 2 * CODEC_STRIDE + threadIdx.x; i < len; i += CODEC_THREADS) {
        const int c = codec_class_of((int64_t)codec_d2(sh, i));
        cls = c > cls ? c : cls;
    }
    if (cls) atomicMax(&s_cls, cls);
    __syncthreads();
    if (threadIdx.x == 0) a.size8[arr][b] = (codec_block_size8(len, s_cls) << 2) | (uint32_t)s_cls;  // size and class
}

// exclusive scan over nb <= CODEC_MAX_BLOCKS entries of (size << 2 | class) -> dir = (offset << 2 | class)
__global__ void __launch_bounds__(1024)
codec_offsets_kernel(CodecArrays a, int64_t nb) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int arr = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += 1024) {
        const int64_t i = base + tid;
        const uint32_t e = i < nb ? a.size8[arr][i] : 0u;
        const uint32_t sz = e >> 2;
        uint32_t incl = sz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += up;
            }
            s_warp[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t before = s_carry + (wid ? s_warp[wid - 1] : 0u) + (incl - sz);
        if (i < nb) a.dir[arr][i] = codec_dir_entry(before, (int)(e & 3u));
        __syncthreads();
        if (tid == 1023) s_carry = before + sz;
        __syncthreads();
    }
    if (tid == 0) a.total8[arr] = (unsigned long long)s_carry;
}

__global__ void __launch_bounds__(CODEC_THREADS)
codec_pack_kernel(CodecArrays a, int64_t cnt, int64_t nb) {
    __shared__ uint64_t sh[CODEC_BLOCK];
    const int arr = blockIdx.y;
    const int64_t b = blockIdx.x;
    const int len = codec_load_block((const uint64_t*)a.values[arr], cnt, b, sh);
    const uint32_t e = a.dir[arr][b];
    const int cls = (int)(e & 3u);
    const uint32_t off8 = e >> 2;
    const uint32_t size8 = codec_block_size8(len, cls);
    if ((uint64_t)off8 + size8 > a.capacity8) return;  // the host sees total8 > capacity8 and takes the plain path
    uint8_t* p = a.payload[arr] + (size_t)off8 * 8;
    uint64_t* p64 = (uint64_t*)p;
    if (threadIdx.x < CODEC_STRIDE) {
        const int i = threadIdx.x;
        p64[i] = i < len ? sh[i] : 0ull;
        p64[CODEC_STRIDE + i] = i + CODEC_STRIDE < len ? sh[i + CODEC_STRIDE] - sh[i] : 0ull;
    }
    // each thread assembles whole 8-byte words of the packed stream: 8 >> cls second differences per word
    const int per = 8 >> cls;
    const int nd2 = len > 2 * CODEC_STRIDE ? len - 2 * CODEC_STRIDE : 0;
    const int nwords = (nd2 + per - 1) / per;
    for (int wd = threadIdx.x; wd < nwords; wd += CODEC_THREADS) {
        uint64_t word = 0;
        for (int j = 0; j < per; ++j) {
            const int i = 2 * CODEC_STRIDE + wd * per + j;
            if (i < len) {
                const uint64_t d2 = codec_d2(sh, i);
                const uint64_t mask = cls == 3 ? ~0ull : ((1ull << (8 << cls)) - 1ull);
                word |= (d2 & mask) << (j * (8 << cls));
            }
        }
        p64[CODEC_HEADER8 + wd] = word;
    }
}

int codec_encode_launch(const CodecArrays& a, int64_t cnt, cudaStream_t st) {
    const int64_t nb = (cnt + CODEC_BLOCK - 1) / CODEC_BLOCK;
    if (nb <= 0) return FGQ_OK;
    if (nb > CODEC_MAX_BLOCKS) {
        set_error("internal: codec chunk too large");
        return FGQ_EINVAL;
    }
    codec_class_kernel<<<dim3((unsigned)nb, 2), CODEC_THREADS, 0, st>>>(a, cnt, nb);
    FGQ_LAUNCH_CHECK();
    codec_offsets_kernel<<<2, 1024, 0, st>>>(a, nb);
    FGQ_LAUNCH_CHECK();
    codec_pack_kernel<<<dim3((unsigned)nb, 2), CODEC_THREADS, 0, st>>>(a, cnt, nb);
    FGQ_LAUNCH_CHECK();
    return FGQ_OK;
}

}  // namespace fgq
