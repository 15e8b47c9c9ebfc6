// gen.cu — device statement of the synthetic column generator; bit-identical to
// baikaldb_b200/datagen.py: integer mixing + exactly rounded double operations only.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../include/bkgpu.h"

namespace {
__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t gen_key(uint64_t seed, uint32_t col, uint32_t k) {
    return mix64(seed + (uint64_t)col * 0xD1B54A32D192ED03ull + (uint64_t)k * 0x8CB92BA72F3D8DD7ull);
}
__device__ inline uint64_t raw64(uint64_t key, int64_t row) { return mix64(key + (uint64_t)(row + 1) * 0x9E3779B97F4A7C15ull); }
__device__ inline double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

struct GenArgs { uint64_t key[4]; uint64_t pkey[4]; int64_t row0, nrows, lo, hi; double scale; int prim, dist, bits; };

__global__ void k_gen(void* dst, GenArgs a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.nrows; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = a.row0 + i;
        int64_t iv = 0; double dv = 0; bool is_d = false;
        switch (a.dist) {
            case 0: iv = (int64_t)(raw64(a.key[0], row) % (uint64_t)(a.hi - a.lo)) + a.lo; break;
            case 1: dv = u01(raw64(a.key[0], row)); is_d = true; break;
            case 2: {
                double s = u01(raw64(a.key[0], row));
                s = __dadd_rn(s, u01(raw64(a.key[1], row)));
                s = __dadd_rn(s, u01(raw64(a.key[2], row)));
                s = __dadd_rn(s, u01(raw64(a.key[3], row)));
                dv = __dmul_rn(__dsub_rn(s, 2.0), a.scale); is_d = true;
            } break;
            case 3: iv = (int64_t)raw64(a.key[0], row); break;
            default: {  // permutation of [0, hi): Feistel + cycle walking
                const int half = a.bits / 2; const uint64_t mask = (1ull << half) - 1ull;
                uint64_t x = (uint64_t)row;
                do {
                    uint64_t l = x >> half, r = x & mask;
                    for (int rd = 0; rd < 4; rd++) { uint64_t f = mix64(r + a.pkey[rd]) & mask; uint64_t nl = r; r = l ^ f; l = nl; }
                    x = (l << half) | r;
                } while (x >= (uint64_t)a.hi);
                iv = (int64_t)x;
            } break;
        }
        if (is_d) ((double*)dst)[i] = dv;
        else if (a.prim == BK_INT64 || a.prim == BK_UINT64) ((int64_t*)dst)[i] = iv;
        else ((int32_t*)dst)[i] = (int32_t)iv;
    }
}
}  // namespace

extern "C" int bkgpu_gen_column(int device, void* dev_dst, int32_t prim_type, int32_t dist, uint64_t seed, uint32_t column_id,
                                int64_t row0, int64_t nrows, int64_t lo, int64_t hi, double scale) {
    if (!dev_dst || nrows < 0 || dist < 0 || dist > 4) return BKGPU_EINVAL;
    if ((dist == 0 || dist == 4) && hi <= lo) return BKGPU_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return BKGPU_ENODEV;
    GenArgs a{};
    for (uint32_t k = 0; k < 4; k++) { a.key[k] = gen_key(seed, column_id, k); a.pkey[k] = gen_key(seed, column_id, 16 + k); }
    a.row0 = row0; a.nrows = nrows; a.lo = lo; a.hi = hi; a.scale = scale; a.prim = prim_type; a.dist = dist;
    int bits = 2; while (((int64_t)1 << bits) < hi) bits++;
    if (hi > 1) { bits = 0; int64_t v = hi - 1; while (v) { bits++; v >>= 1; } if (bits < 2) bits = 2; }
    bits += bits & 1; a.bits = bits;
    if (nrows == 0) return BKGPU_OK;
    int grid = (int)((nrows + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    k_gen<<<grid, 256>>>(dev_dst, a);
    if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return BKGPU_ENODEV;
    return BKGPU_OK;
}
