// sw_score_long.cuh -- score-only local alignment for long streamed sequences (rows > 1024), sm_100a.
//
// Same recurrence and packed DPX arithmetic as sw_score.cuh (zoe: sw_simd_score,
// src/alignment/sw/striped.rs:65-142), for the ONT-read-vs-genome shape (BASELINE config 4): one warp
// owns a task (two sequences packed in the 16-bit halves), the rows are cut into chunks of up to
// 32 lanes x K rows that the warp sweeps one after the other over all columns; the H / E values leaving
// a chunk's last row are parked in a per-warp boundary row in global memory (8 bytes per column,
// L2-resident) and picked up as the top boundary of the next chunk:
//     * bottom boundary: lane 31 stores (H, E^) of its last row, one 8-byte store per column;
//     * top boundary: all lanes prefetch the next 32 columns with one coalesced 256-byte load, and
//       lane 0 receives its column's entry by __shfl_sync -- the load latency never sits on the systolic
//       critical path.
// The last (partial) chunk only executes the 4-row blocks it needs (a warp-uniform branch), so the
// padding waste is < 128 rows per task.  Tasks are sorted longest-first on the host and handed out through
// an atomic queue, so warps stay balanced although lengths vary 5x.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "sw_score.cuh"

namespace zoe_cuda {

struct LongParams {
    ScoreParams s;           // s.task_ids is required (length-sorted sequence ids), s.n_rseq = list length
    uint2 *boundary;         // [warp slots][max_L]
    uint32_t max_L;
    unsigned int *queue;     // dynamic task counter (zeroed before launch)
};

template <int K, bool PACKED>
__global__ void __launch_bounds__(512) sw_score_long_kernel(const LongParams lp) {
    using O = Ops<PACKED>;
    const ScoreParams &p = lp.s;
    constexpr int G = 32;
    constexpr int K4 = K / 4;
    static_assert(K % 4 == 0, "K must be a multiple of 4");
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) uint8_t smem[];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int warps_per_block = blockDim.x >> 5;

    const int tab_bytes = p.n_csym * K4 * G * 16;
    uint4 *tab = reinterpret_cast<uint4 *>(smem + (size_t)warp * tab_bytes);
    uint8_t *s_lut = smem + (size_t)warps_per_block * tab_bytes;
    int8_t *s_wk = reinterpret_cast<int8_t *>(s_lut + 256);
    uint8_t *s_cc = reinterpret_cast<uint8_t *>(s_wk) + ((p.n_csym * p.S + 15) & ~15);
    for (int i = tid; i < 256; i += blockDim.x) s_lut[i] = p.lut[i];
    for (int i = tid; i < p.n_csym * p.S; i += blockDim.x) s_wk[i] = p.wk[i];
    if (p.cols_in_smem)
        for (uint32_t i = tid; i < p.ccodes_bytes; i += blockDim.x) s_cc[i] = p.ccodes[i];
    __syncthreads();
    const uint8_t *cc = p.cols_in_smem ? s_cc : p.ccodes;

    const uint32_t go_s = O::splat(p.go), neg_ge = O::splat(-p.ge);
    uint2 *bnd = lp.boundary + (size_t)(blockIdx.x * warps_per_block + warp) * lp.max_L;

    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(lp.queue, 1u);
        task = __shfl_sync(FULL, task, 0);
        if (task >= p.n_tasks) break;

        uint32_t id_lo, id_hi = 0xffffffffu;
        if (PACKED) {
            id_lo = p.task_ids[2 * task];
            if (2 * task + 1 < p.n_rseq) id_hi = p.task_ids[2 * task + 1];
        } else {
            id_lo = p.task_ids[task];
        }
        const uint64_t off_lo = p.roff[id_lo];
        const int len_lo = (int)(p.roff[id_lo + 1] - off_lo);
        uint64_t off_hi = 0;
        int len_hi = 0;
        if (PACKED && id_hi != 0xffffffffu) {
            off_hi = p.roff[id_hi];
            len_hi = (int)(p.roff[id_hi + 1] - off_hi);
        }
        const int nmax = max(len_lo, len_hi);
        const int R = G * K;
        const int nchunks = max(1, (nmax + R - 1) / R);

        for (uint32_t cj = 0; cj < p.n_cseq; ++cj) {
            const uint32_t c0 = p.coff[cj];
            const int L = (int)(p.coff[cj + 1] - c0);
            const uint8_t *cs = cc + c0;
            uint32_t best = 0;

            for (int ch = 0; ch < nchunks; ++ch) {
                const int row0 = ch * R;
                const int rows_left = nmax - row0;
                const int k4_eff = min(K4, max(1, (rows_left + 4 * G - 1) / (4 * G)));
                const int k_eff = 4 * k4_eff;
                const bool has_top = ch > 0, has_bottom = ch + 1 < nchunks;

                // ---- score table of this chunk: lane l owns rows row0 + l*k_eff + [0, k_eff) ----
                __syncwarp();
                for (int s = 0; s < p.n_csym; ++s) {
                    const int8_t *wrow = s_wk + s * p.S;
                    for (int i4 = 0; i4 < k4_eff; ++i4) {
                        uint32_t w[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int r = row0 + lane * k_eff + i4 * 4 + q;
                            int wl = kPadWeight, wh = kPadWeight;
                            if (r < len_lo) wl = wrow[s_lut[p.rseq[off_lo + r]]];
                            if (PACKED && r < len_hi) wh = wrow[s_lut[p.rseq[off_hi + r]]];
                            w[q] = PACKED ? ((uint32_t)(wl & 0xffff) | ((uint32_t)(wh & 0xffff) << 16)) : (uint32_t)wl;
                        }
                        tab[(s * K4 + i4) * G + lane] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                __syncwarp();

                uint32_t Hrow[K], Frow[K];
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    Hrow[i] = 0;
                    Frow[i] = 0;
                }
                uint32_t h_last = 0, e_out = 0, h_up_prev = 0;
                uint2 cur = make_uint2(0, 0), nxt = make_uint2(0, 0);
                if (has_top && lane < L) cur = __ldcg(bnd + lane);
                const int nsteps = L + G - 1;

                for (int step = 0; step < nsteps; ++step) {
                    if (has_top && (step & 31) == 0) {
                        const int idx = step + 32 + lane;
                        nxt = (idx < L) ? __ldcg(bnd + idx) : make_uint2(0, 0);
                    }
                    uint32_t h_in = __shfl_up_sync(FULL, h_last, 1);
                    uint32_t e_in = __shfl_up_sync(FULL, e_out, 1);
                    const uint32_t h_top = __shfl_sync(FULL, cur.x, step & 31);
                    const uint32_t e_top = __shfl_sync(FULL, cur.y, step & 31);
                    if (lane == 0) {
                        h_in = has_top ? h_top : 0;
                        e_in = has_top ? e_top : 0;
                    }
                    if ((step & 31) == 31) cur = nxt;
                    const int j = step - lane;
                    if (j >= 0 && j < L) {
                        const int s = cs[j];
                        const uint4 *tp = tab + (size_t)s * (K4 * G) + lane;
                        uint32_t diag = h_up_prev;
                        uint32_t E = e_in;
#pragma unroll
                        for (int i4 = 0; i4 < K4; ++i4) {
                            if (i4 < k4_eff) {
                                const uint4 w4 = tp[i4 * G];
                                const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
                                uint32_t hp = 0;
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const int i = i4 * 4 + q;
                                    uint32_t x = O::max3(E, Frow[i], go_s) - go_s;
                                    uint32_t H = O::addmax(diag, w[q], x);
                                    diag = Hrow[i];
                                    E = O::addmax(E, neg_ge, H);
                                    Frow[i] = O::addmax(Frow[i], neg_ge, H);
                                    Hrow[i] = H;
                                    if (q & 1)
                                        best = O::max3(best, H, hp);
                                    else
                                        hp = H;
                                }
                                h_last = Hrow[i4 * 4 + 3];
                            }
                        }
                        e_out = E;
                        if (has_bottom && lane == G - 1) bnd[j] = make_uint2(h_last, e_out);
                    }
                    h_up_prev = h_in;
                }
                __syncwarp();
            }

#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) best = O::max2(best, __shfl_xor_sync(FULL, best, d));
            if (lane == 0) {
                if (PACKED) {
                    int b_lo = (int)(int16_t)(best & 0xffff), b_hi = (int)(int16_t)(best >> 16);
                    p.best[(size_t)id_lo * p.n_cseq + cj] = (b_lo >= p.ovf_thresh) ? -1 : b_lo;
                    if (id_hi != 0xffffffffu) p.best[(size_t)id_hi * p.n_cseq + cj] = (b_hi >= p.ovf_thresh) ? -1 : b_hi;
                } else {
                    p.best[(size_t)id_lo * p.n_cseq + cj] = (int)best;
                }
            }
        }
    }
}

}  // namespace zoe_cuda
