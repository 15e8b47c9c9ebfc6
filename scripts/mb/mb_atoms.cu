// Microbenchmark of the lean aggregate kernel's DRAIN alone (no HBM traffic): what one pass of 32 queue entries costs in SM cycles
// under each way of updating {count, sumA, sumB} of a random group slot in the CTA's shared table.  Same launch shape as
// k_agg_group_lean (148 x 640 threads, one CTA per SM).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_atoms mb_atoms.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t lds64(uint32_t a) { uint64_t v; asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(uint32_t a, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void reds_inc32(uint32_t a) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); }
__device__ __forceinline__ void reds_add32(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t atoms_add32(uint32_t a, uint32_t v) { uint32_t o; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }
__device__ __forceinline__ void lds128(uint32_t a, uint64_t& lo, uint64_t& hi) { asm volatile("ld.volatile.shared.v2.u64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(a)); }
__device__ __forceinline__ void sts128(uint32_t a, uint64_t lo, uint64_t hi) { asm volatile("st.shared.v2.u64 [%0], {%1,%2};" ::"r"(a), "l"(lo), "l"(hi) : "memory"); }
__device__ __forceinline__ void atoms_cas128(uint32_t a, uint64_t c0, uint64_t c1, uint64_t n0, uint64_t n1, uint64_t& p0, uint64_t& p1) {
    asm volatile("{\n .reg .b128 c, n, p;\n mov.b128 c, {%2, %3};\n mov.b128 n, {%4, %5};\n atom.shared.cas.b128 p, [%6], c, n;\n mov.b128 {%0, %1}, p;\n}"
                 : "=l"(p0), "=l"(p1) : "l"(c0), "l"(c1), "l"(n0), "l"(n1), "r"(a) : "memory");
}
__device__ __forceinline__ uint64_t atoms_cas64(uint32_t a, uint64_t cmp, uint64_t nw) {
    uint64_t o; asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(o) : "r"(a), "l"(cmp), "l"(nw) : "memory"); return o;
}
__device__ __forceinline__ double bits_f64(uint64_t b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ uint64_t f64_bits(double d) { return (uint64_t)__double_as_longlong(d); }
__device__ __forceinline__ void add_f64_cas(uint32_t a, double v) {
    uint64_t cur = lds64(a);
    for (;;) { const uint64_t nw = f64_bits(bits_f64(cur) + v); const uint64_t prev = atoms_cas64(a, cur, nw); if (prev == cur) break; cur = prev; }
}
constexpr int CAP = 2048, THREADS = 640, WARPS = THREADS / 32, QN = 64;
constexpr uint64_t EMPTY = ~0ull;
// shared layout: keys[CAP] u64 | cnt/pad [CAP] 16 B | sums [CAP] 16 B | limbs [CAP] 32 B | queues
template <int V>
__global__ void __launch_bounds__(THREADS, 1) k(int reps, int ngroups, double scale, double scale1, unsigned long long* sink) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* keys = (uint64_t*)sm;
    const uint32_t keys_a = smem_addr(sm), cnt_a = keys_a + CAP * 8, sum_a = cnt_a + CAP * 16, limb_a = sum_a + CAP * 16, q_a = limb_a + CAP * 32 + warp * (QN * 24);
    for (int i = threadIdx.x; i < CAP; i += THREADS) keys[i] = EMPTY;
    for (int i = threadIdx.x; i < CAP * 8; i += THREADS) ((uint64_t*)(sm + CAP * 8))[i] = 0;
    __syncthreads();
    // populate the key table with every group (as the real kernel's table is after its first iterations)
    for (int g = threadIdx.x; g < ngroups; g += THREADS) {
        uint32_t slot = ((uint32_t)g * 0x9E3779B1u) >> 21;
        for (;;) { const uint64_t o = atoms_cas64(keys_a + slot * 8, EMPTY, (uint64_t)g); if (o == EMPTY || o == (uint64_t)g) break; slot = (slot + 1) & (CAP - 1); }
    }
    uint64_t x = (blockIdx.x * 977u + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    for (int e = lane; e < QN; e += 32) {
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        sts64(q_a + e * 8, x % (uint64_t)ngroups);
        sts64(q_a + (QN + e) * 8, f64_bits((double)(x >> 11) * (1.0 / 9007199254740992.0)));
        sts64(q_a + (2 * QN + e) * 8, f64_bits(((double)(int64_t)(x * 31) * (1.0 / 9223372036854775808.0)) * 3000.0));
    }
    __syncthreads();
    uint32_t bump = 0;
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
#pragma unroll 1
        for (int e0 = 0; e0 < QN; e0 += 32) {
            const int e = e0 + lane;
            uint64_t k0 = lds64(q_a + e * 8) + bump;
            if (k0 >= (uint64_t)ngroups) k0 -= ngroups;
            const uint64_t v0 = lds64(q_a + (QN + e) * 8), v1 = lds64(q_a + (2 * QN + e) * 8);
            uint32_t slot;
            if (V == 6 || V == 7) slot = (uint32_t)k0;
            else {
                slot = ((uint32_t)k0 * 0x9E3779B1u) >> 21;
                for (;;) { const uint64_t kk = lds64(keys_a + slot * 8); if (kk == k0) break; slot = (slot + 1) & (CAP - 1); }
            }
            if (V >= 1 && V != 11) reds_inc32(cnt_a + slot * 16);
            if (V == 2 || V == 6) {
                const uint32_t addr = sum_a + slot * 16; uint64_t c0, c1; lds128(addr, c0, c1);
                for (;;) { uint64_t p0, p1; atoms_cas128(addr, c0, c1, f64_bits(bits_f64(c0) + bits_f64(v0)), f64_bits(bits_f64(c1) + bits_f64(v1)), p0, p1);
                    if (p0 == c0 && p1 == c1) break; c0 = p0; c1 = p1; }
            }
            if (V == 3 || V == 7 || V == 9) {   // 64-bit fixed point, two 32-bit limbs per sum: {lo0, hi0, lo1, hi1} in one 16-byte word
                const uint32_t addr = sum_a + slot * 16;
                const long long f0 = __double2ll_rn(bits_f64(v0) * scale), f1 = __double2ll_rn(bits_f64(v1) * scale);
                const uint32_t l0 = (uint32_t)f0, h0 = (uint32_t)((uint64_t)f0 >> 32), l1 = (uint32_t)f1, h1 = (uint32_t)((uint64_t)f1 >> 32);
                if (V == 9) { reds_add32(addr, l0); reds_add32(addr + 8, l1); reds_add32(addr + 4, h0); reds_add32(addr + 12, h1); }
                else {
                    const uint32_t o0 = atoms_add32(addr, l0), o1 = atoms_add32(addr + 8, l1);
                    reds_add32(addr + 4, h0 + ((o0 + l0) < o0 ? 1u : 0u));
                    reds_add32(addr + 12, h1 + ((o1 + l1) < o1 ? 1u : 0u));
                }
            }
            if (V == 4) {   // 96-bit fixed point, three limbs per sum: {lo, mid, hi, -} x 2 in a 32-byte record
                const uint32_t addr = limb_a + slot * 32;
                const double a0 = bits_f64(v0) * scale, a1 = bits_f64(v1) * scale;           // |a| < 2^73
                const double t0 = rint(a0 * (1.0 / 4294967296.0)), t1 = rint(a1 * (1.0 / 4294967296.0));   // upper 64 bits (exact split)
                const long long u0 = __double2ll_rn(t0), u1 = __double2ll_rn(t1);
                const long long w0 = __double2ll_rn(a0 - t0 * 4294967296.0), w1 = __double2ll_rn(a1 - t1 * 4294967296.0);   // |w| <= 2^31
                // value = u * 2^32 + w  ->  limbs: lo = (u32)w, mid = (u32)u + signext(w) , hi = (u32)(u>>32) + ...
                const uint32_t lo0 = (uint32_t)w0, lo1 = (uint32_t)w1;
                const uint64_t m0 = (uint64_t)u0 + (uint64_t)(w0 >> 32), m1 = (uint64_t)u1 + (uint64_t)(w1 >> 32);
                const uint32_t oa = atoms_add32(addr, lo0), ob = atoms_add32(addr + 16, lo1);
                const uint64_t n0 = m0 + ((oa + lo0) < oa ? 1u : 0u), n1 = m1 + ((ob + lo1) < ob ? 1u : 0u);
                const uint32_t ma = (uint32_t)n0, mb = (uint32_t)n1;
                const uint32_t pa = atoms_add32(addr + 4, ma), pb = atoms_add32(addr + 20, mb);
                reds_add32(addr + 8, (uint32_t)(n0 >> 32) + ((pa + ma) < pa ? 1u : 0u));
                reds_add32(addr + 24, (uint32_t)(n1 >> 32) + ((pb + mb) < pb ? 1u : 0u));
            }
            if (V == 10 || V == 11) {   // planned layout: one u32 array per limb (4-byte stride: 32 banks), main / fine / CAS classification
                if (V == 11) reds_inc32(limb_a + 6 * CAP * 4 + slot * 4);
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    const double xv = bits_f64(s ? v1 : v0);
                    const double y = xv * (s ? scale1 : scale);
                    const double ay = fabs(y);
                    const uint32_t mid_a = limb_a + (s * 3 + 0) * CAP * 4 + slot * 4, hi_a = limb_a + (s * 3 + 1) * CAP * 4 + slot * 4, ext_a = limb_a + (s * 3 + 2) * CAP * 4 + slot * 4;
                    if (ay >= 67108864.0 && ay < 2199023255552.0) {
                        const long long f = __double2ll_rn(y);
                        const uint32_t lo = (uint32_t)f, hi = (uint32_t)((uint64_t)f >> 32);
                        const uint32_t o = atoms_add32(mid_a, lo);
                        const uint32_t h = hi + ((o + lo) < o ? 1u : 0u);
                        if (h) reds_add32(hi_a, h);
                    } else if (xv == 0.0) {
                    } else if (ay >= 0.015625 && ay < 67108864.0) {
                        const long long f = __double2ll_rn(y * 4294967296.0);
                        const uint32_t e = (uint32_t)f;
                        const uint32_t o = atoms_add32(ext_a, e);
                        const uint64_t t = (uint64_t)(f >> 32) + ((o + e) < o ? 1u : 0u);
                        const uint32_t m = (uint32_t)t;
                        const uint32_t o2 = atoms_add32(mid_a, m);
                        const uint32_t h = (uint32_t)(t >> 32) + ((o2 + m) < o2 ? 1u : 0u);
                        if (h) reds_add32(hi_a, h);
                    } else add_f64_cas(sum_a + slot * 16 + s * 8, xv);
                }
            }
            if (V == 5) { add_f64_cas(sum_a + slot * 16, bits_f64(v0)); add_f64_cas(sum_a + slot * 16 + 8, bits_f64(v1)); }
            if (V == 8) { const uint32_t addr = sum_a + slot * 16; uint64_t c0, c1; lds128(addr, c0, c1);
                sts128(addr, f64_bits(bits_f64(c0) + bits_f64(v0)), f64_bits(bits_f64(c1) + bits_f64(v1))); }
        }
        bump = (bump + 7) % (uint32_t)ngroups;
        __syncwarp();
    }
    __syncthreads();
    unsigned long long acc = 0;
    for (int i = threadIdx.x; i < CAP * 8; i += THREADS) acc += ((uint64_t*)(sm + CAP * 8))[i];
    if (acc == 0x1234567) sink[0] = acc;
}
template <int V> static void run(const char* name, int reps, int ngroups, double clk_ghz, unsigned long long* sink) {
    const size_t smem = CAP * (8 + 16 + 16 + 32) + WARPS * QN * 24;
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<V><<<148, THREADS, smem>>>(reps / 10, ngroups, 274877906944.0, 67108864.0, sink);
    cudaEventRecord(a); k<V><<<148, THREADS, smem>>>(reps, ngroups, 274877906944.0, 67108864.0, sink); cudaEventRecord(b);
    cudaError_t e = cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double passes = (double)reps * (QN / 32) * WARPS;   // per SM
    printf("%-44s groups=%5d  %8.3f ms  %7.1f SM-cycles per 32-entry pass  (%s)\n", name, ngroups, ms, ms * 1e-3 * clk_ghz * 1e9 / passes, cudaGetErrorString(e));
}
int main() {
    unsigned long long* sink; cudaMalloc(&sink, 8);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk * 1e-6;
    printf("SM clock (max) %.3f GHz; lean kernel today: ~95 SM-cycles per pass all-in (1.0M cycles / 10558 passes per SM)\n", ghz);
    for (int ng : {1000, 100}) {
        const int R = 4000;
        run<0>("v0 queue read + key probe only", R, ng, ghz, sink);
        run<1>("v1 + RED.u32 count", R, ng, ghz, sink);
        run<2>("v2 + count + LDS.128/CAS.128 (today)", R, ng, ghz, sink);
        run<5>("v5 + count + 2 x (LDS.64/CAS.64)", R, ng, ghz, sink);
        run<3>("v3 + count + fx64: 2 x (ATOMS.ADD ret + RED)", R, ng, ghz, sink);
        run<9>("v9 + count + 4 x RED (no carry; bound)", R, ng, ghz, sink);
        run<4>("v4 + count + fx96: 2 x (2 ATOMS.ADD ret + RED)", R, ng, ghz, sink);
        run<8>("v8 + count + LDS.128/STS.128 (racy; bound)", R, ng, ghz, sink);
        run<6>("v6 today, identity slot (no probe)", R, ng, ghz, sink);
        run<7>("v7 fx64, identity slot (no probe)", R, ng, ghz, sink);
        run<10>("v10 fx64 SoA limbs + classification", R, ng, ghz, sink);
        run<11>("v11 fx64 SoA limbs, SoA count", R, ng, ghz, sink);
    }
    return 0;
}
