// ubench_scan.cu -- micro-benchmark of pass A's steady loop (sw_align_scan_kernel, sw_align_win.cuh): does splitting a
// lane's K rows into two halves that work on neighbouring columns (two independent dependency chains per thread, the
// same instructions) raise the ALU-pipe utilisation?  Synthetic data, no results kept beyond a checksum.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_scan scripts/ubench_scan.cu && ./ubench_scan
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define FULL 0xffffffffu
__device__ __forceinline__ uint32_t addmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
__device__ __forceinline__ uint32_t min2(uint32_t a, uint32_t b) { return __vmins2(a, b); }

constexpr int G = 8;

// KA == K: one chain (today's kernel).  KA < K: rows [0, KA) sweep column s - 2*lig, rows [KA, K) column s - 2*lig - 1.
// FEAT bit 0: gap penalties are kernel arguments (the real kernel; otherwise immediates)
// FEAT bit 1: checkpoint stores (2K + 2 words per thread every 64 steps)
// FEAT bit 2: column codes vary (otherwise every column reads the same table row)
template <int K, int KA, int THREADS, int FEAT = 0>
__global__ void __launch_bounds__(THREADS) scan_loop(const uint8_t *cols, int L, int trips, uint32_t *out, uint32_t go_arg,
                                                     uint32_t ge_arg, uint32_t *ckpt) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int K4 = (K + 3) / 4;
    constexpr int NSYM = 4;
    constexpr int TAB = NSYM * K4 * G * 16;
    const int tid = threadIdx.x, lig = tid % G, grp = tid / G, ngrp = blockDim.x / G;
    uint8_t *s_cc = smem + (size_t)ngrp * TAB;
    for (int i = tid; i < L + 64; i += blockDim.x) s_cc[i] = (FEAT & 4) ? (uint8_t)((i * 2654435761u) >> 30) : (cols[i % L] & 3);
    for (int i = tid; i < ngrp * TAB / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(smem)[i] = ((i * 2654435761u) >> 28) % 3 == 0 ? 0x00020002u : 0xfffbfffbu;
    __syncthreads();
    uint32_t tab_lane_off = (uint32_t)grp * TAB + (uint32_t)lig * 16u;
    uint32_t go_s = (FEAT & 1) ? go_arg : 0x000a000au, neg_ge = (FEAT & 1) ? ge_arg : 0xffffffffu, one_s = 0x00010001u;
    uint32_t nz_lane = lig != 0;
    // FEAT bit 3: the penalties stay kernel arguments the compiler may read straight from the constant bank
    if (FEAT & 8) {
        go_s = go_arg;
        neg_ge = ge_arg;
        asm volatile("" : "+r"(tab_lane_off), "+r"(nz_lane));
    } else {
        asm volatile("" : "+r"(go_s), "+r"(neg_ge), "+r"(tab_lane_off), "+r"(nz_lane));
    }
    const uint32_t neg_go = (FEAT & 8) ? 0u - go_arg : 0u;
    uint32_t acc = 0;
    for (int trip = 0; trip < trips; ++trip) {
        uint32_t H[2][K], F[K];
#pragma unroll
        for (int i = 0; i < K; ++i) H[0][i] = H[1][i] = F[i] = 0;
        uint32_t h_last = 0, e_out = 0, h_up_prev = 0;      // half A (or the whole lane)
        uint32_t hA_prev = 0, hB_last = 0, eB_out = 0;       // split: A's h_last one step earlier; half B's outputs
        uint32_t bestp = 0, firstp = 0, lastp = 0, cm = 0, hp_pend = 0;
        const uint8_t *cs = s_cc;

        auto rows = [&](auto parity, auto lo, auto hi, const int j, uint32_t diag, uint32_t E, uint32_t &h_out, uint32_t &e_o) {
            constexpr int PO = decltype(parity)::value, PN = 1 - PO;
            constexpr int R0 = decltype(lo)::value, R1 = decltype(hi)::value;
            const uint4 *tp = reinterpret_cast<const uint4 *>(smem + tab_lane_off + (uint32_t)cs[j] * (uint32_t)(K4 * G * 16));
            uint32_t hp = hp_pend;
#pragma unroll
            for (int i4 = R0 / 4; i4 < (R1 + 3) / 4; ++i4) {
                const uint4 w4 = tp[i4 * G];
                const uint32_t wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = i4 * 4 + q;
                    if (i >= R0 && i < R1) {
                        const uint32_t x = (FEAT & 8) ? max3(E, F[i], go_s) + neg_go : max3(E, F[i], go_s) - go_s;
                        const uint32_t Hn = addmax(diag, wv[q], x);
                        diag = H[PO][i];
                        E = addmax(E, neg_ge, Hn);
                        F[i] = addmax(F[i], neg_ge, Hn);
                        H[PN][i] = Hn;
                        if ((i + PO) & 1)
                            cm = max3(cm, Hn, hp);
                        else
                            hp = Hn;
                    }
                }
            }
            hp_pend = hp;
            h_out = H[PN][R1 - 1];
            e_o = E;
        };
        auto bookkeeping = [&](const uint32_t s_even) {
            const uint32_t pp = (s_even >> 1) * 0x00010001u;
            const uint32_t m = max2(cm, bestp);
            const uint32_t inc = min2(m - bestp, one_s) * 0xffffu;
            const uint32_t ge = (one_s - min2(m - cm, one_s)) * 0xffffu;
            firstp = (firstp & ~inc) | (pp & inc);
            lastp = (lastp & ~ge) | (pp & ge);
            bestp = m;
            cm = 0;
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        using R0 = std::integral_constant<int, 0>;
        using RA = std::integral_constant<int, KA>;
        using RK = std::integral_constant<int, K>;
        for (int s = 2 * G; s + 1 < L; s += 2) {
            if ((FEAT & 2) && (s & 63) == 0) {
                uint32_t *dst = ckpt + ((size_t)(blockIdx.x * (THREADS / G) + grp) * 32 + (s >> 6)) * ((2 * K + 2) * G) + lig;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    dst[i * G] = H[0][i];
                    dst[(K + i) * G] = F[i];
                }
                dst[(2 * K) * G] = e_out;
                dst[(2 * K + 1) * G] = h_up_prev;
            }
            if (KA == K) {
                {
                    const uint32_t h_in = __shfl_up_sync(FULL, h_last, 1, G) * nz_lane;
                    const uint32_t e_in = __shfl_up_sync(FULL, e_out, 1, G) * nz_lane;
                    rows(I0{}, R0{}, RK{}, s - lig, h_up_prev, e_in, h_last, e_out);
                    h_up_prev = h_in;
                }
                {
                    const uint32_t h_in = __shfl_up_sync(FULL, h_last, 1, G) * nz_lane;
                    const uint32_t e_in = __shfl_up_sync(FULL, e_out, 1, G) * nz_lane;
                    rows(I1{}, R0{}, RK{}, s + 1 - lig, h_up_prev, e_in, h_last, e_out);
                    h_up_prev = h_in;
                }
            } else {
                {
                    const uint32_t h_in = __shfl_up_sync(FULL, hB_last, 1, G) * nz_lane;
                    const uint32_t e_in = __shfl_up_sync(FULL, eB_out, 1, G) * nz_lane;
                    const uint32_t hA = h_last, eA = e_out;  // A's outputs of the previous step: column s - 2*lig - 1
                    rows(I0{}, RA{}, RK{}, s - 2 * lig - 1, hA_prev, eA, hB_last, eB_out);
                    rows(I0{}, R0{}, RA{}, s - 2 * lig, h_up_prev, e_in, h_last, e_out);
                    hA_prev = hA;
                    h_up_prev = h_in;
                }
                {
                    const uint32_t h_in = __shfl_up_sync(FULL, hB_last, 1, G) * nz_lane;
                    const uint32_t e_in = __shfl_up_sync(FULL, eB_out, 1, G) * nz_lane;
                    const uint32_t hA = h_last, eA = e_out;
                    rows(I1{}, RA{}, RK{}, s - 2 * lig, hA_prev, eA, hB_last, eB_out);
                    rows(I1{}, R0{}, RA{}, s + 1 - 2 * lig, h_up_prev, e_in, h_last, e_out);
                    hA_prev = hA;
                    h_up_prev = h_in;
                }
            }
            bookkeeping((uint32_t)s);
        }
        acc += bestp + firstp + lastp + H[0][K - 1] + F[0];
    }
    out[blockIdx.x * blockDim.x + tid] = acc;
}

template <int K, int KA, int THREADS, int FEAT = 0>
void run(const char *name, const uint8_t *d_cols, uint32_t *d_out, int L, int trips) {
    auto fn = scan_loop<K, KA, THREADS, FEAT>;
    static uint32_t *d_ck = nullptr;
    if (!d_ck) cudaMalloc(&d_ck, (size_t)148 * 128 * 32 * 64 * 8 * 4);
    constexpr int K4 = (K + 3) / 4;
    const size_t smem = (size_t)(THREADS / G) * 4 * K4 * G * 16 + L + 64;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        printf("%-28s smem %zu does not fit\n", name, smem);
        cudaGetLastError();
        return;
    }
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, THREADS, smem);
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, fn);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        fn<<<148, THREADS, smem>>>(d_cols, L, trips, d_out, 0x000a000au, 0xffffffffu, d_ck);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    cudaError_t e = cudaGetLastError();
    const double pairs = 148.0 * THREADS * K * (double)((L - 2 * G) / 2 * 2) * trips;  // packed cell pairs
    printf("%-28s regs %3d spill %3zu occ %d  %8.3f ms  %7.1f GCUPS  %s\n", name, fa.numRegs, (size_t)fa.localSizeBytes, nb, best,
           2.0 * pairs / (best * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const int L = 1704, trips = 40;
    uint8_t *d_cols;
    uint32_t *d_out;
    cudaMalloc(&d_cols, L);
    cudaMemset(d_cols, 1, L);
    cudaMalloc(&d_out, 148 * 1024 * 4);
    run<19, 19, 640>("K19 one chain, 640 thr", d_cols, d_out, L, trips);
    run<19, 19, 640, 1>("K19 640 +runtime gaps", d_cols, d_out, L, trips);
    run<19, 19, 640, 2>("K19 640 +checkpoints", d_cols, d_out, L, trips);
    run<19, 19, 640, 4>("K19 640 +varying codes", d_cols, d_out, L, trips);
    run<19, 19, 640, 7>("K19 640 +all three", d_cols, d_out, L, trips);
    run<19, 19, 640, 8>("K19 640 const-bank gaps", d_cols, d_out, L, trips);
    run<19, 19, 640, 14>("K19 640 cbank+ckpt+codes", d_cols, d_out, L, trips);
    run<19, 12, 640>("K19 split 12/7, 640 thr", d_cols, d_out, L, trips);
    run<19, 8, 640>("K19 split 8/11, 640 thr", d_cols, d_out, L, trips);
    run<20, 20, 640>("K20 one chain, 640 thr", d_cols, d_out, L, trips);
    run<20, 12, 640>("K20 split 12/8, 640 thr", d_cols, d_out, L, trips);
    run<19, 19, 512>("K19 one chain, 512 thr", d_cols, d_out, L, trips);
    run<19, 12, 512>("K19 split 12/7, 512 thr", d_cols, d_out, L, trips);
    run<20, 12, 512>("K20 split 12/8, 512 thr", d_cols, d_out, L, trips);
    run<24, 24, 512>("K24 one chain, 512 thr", d_cols, d_out, L, trips);
    run<24, 12, 512>("K24 split 12/12, 512 thr", d_cols, d_out, L, trips);
    return 0;
}
