// sw_score_rows.cuh -- score-only local alignment with the PROFILED sequence register-resident (sm_100a).
//
// Same contract as sw_score.cuh (zoe: sw_simd_score, src/alignment/sw/striped.rs:65-142, with the escalation of
// src/alignment/profile_set.rs:71-78).  sw_score_kernel keeps the streamed sequences in registers and needs one
// score table PER TASK (column symbols x rows x 4 B): 2.5 kB for a 150-nt read pair, but 25 kB for a pair of 300-aa
// queries against a 20-letter target, which leaves 4 warps per SM (BASELINE config 5: 0.29 of the roofline).
// The local-alignment score is symmetric under transposition, so here the roles are swapped:
//
//   * rows     = the profiled sequence (<= G*K = 1024 residues), K rows per lane in registers, the same rows for
//                every task, hence ONE score table per CTA: tab[streamed symbol][row/4][lane] = uint4 of four row
//                weights, each zero-extended to 32 bits;
//   * columns  = the streamed sequences; a task is two of them (the two 16-bit halves), whose column symbols differ,
//                so the packed weight is assembled per row as  w = w_hi * 65536 + w_lo  (one IMAD on the FMA pipe,
//                which has slack: the ALU pipe bounds the kernel);
//   * a group (warp) sweeps TWO tasks at once as independent dependency chains (the table is shared, so the second
//                chain costs registers only).
//
// Per packed cell pair: 4.5 DPX/ALU + 2 FMA-pipe (x - go, weight assembly) + 0.5 LDS.128.  The sweep is as short as a
// streamed sequence (300 columns in config 5), so the lane skew (G-1 steps of ramp and tail) is a real cost here:
// skew 1 is used (the shuffle latency is covered by the second chain).
//
// Columns past the end of the shorter sequence of a task use a pad symbol whose weights are kPadWeight: such cells
// can only be reached through a gap from a real cell and therefore never exceed the true best (same argument as the
// padding rows of sw_score_kernel).
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "sw_score.cuh"

namespace zoe_cuda {

struct RowsParams {
    ScoreParams s;        // rseq/roff = streamed (columns here), ccodes/coff/wk = profiled (rows here)
    uint32_t cj;          // which profiled sequence this launch scores
    uint32_t max_rlen;    // longest streamed sequence (sizes the per-warp staging area)
};

// bytes of the CTA-wide table: (S + 1 pad symbol) x K4 x G uint4
__host__ __device__ inline size_t rows_tab_bytes(int S, int G, int K) { return (size_t)(S + 1) * ((K + 3) / 4) * G * 16; }
// per-warp staging: 2 chains x 2 halves x max_rlen streamed symbol indices (uint16)
__host__ __device__ inline size_t rows_stage_bytes(uint32_t max_rlen) { return (size_t)4 * ((max_rlen + 7) & ~7u) * 2; }

template <int G, int K>
__global__ void __launch_bounds__(K > 20 ? 256 : 384) sw_score_rows_kernel(const RowsParams rp) {
    constexpr bool PACKED = true;
    using O = Ops<true>;
    const ScoreParams &p = rp.s;
    constexpr int K4 = (K + 3) / 4;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(G == 32, "one task pair per warp");
    extern __shared__ __align__(16) uint8_t smem[];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int lig = lane;

    // ---- shared memory: [table (S+1) x K4 x G uint4][per-warp column offsets] ----
    const uint32_t sym_stride = (uint32_t)(K4 * G * 16);  // bytes per streamed symbol
    const size_t tab_bytes = rows_tab_bytes(p.S, G, K);
    const uint32_t stage_len = (rp.max_rlen + 7) & ~7u;
    uint16_t *stage = reinterpret_cast<uint16_t *>(smem + tab_bytes) + (size_t)warp * 4 * stage_len;

    // ---- build the table once per CTA: tab[q][i4][l].e = weight(streamed symbol q, profiled row l*K + 4*i4 + e) ----
    const uint32_t c0 = p.coff[rp.cj];
    const int Lrows = (int)(p.coff[rp.cj + 1] - c0);
    for (int idx = tid; idx < (p.S + 1) * K4 * G * 4; idx += blockDim.x) {
        const int e = idx & 3, l = (idx >> 2) % G, i4 = ((idx >> 2) / G) % K4, q = (idx >> 2) / (G * K4);
        const int i = i4 * 4 + e, r = l * K + i;
        int w = kPadWeight;
        if (q < p.S && i < K && r < Lrows) w = (int)p.wk[(size_t)p.ccodes[c0 + r] * p.S + q];
        reinterpret_cast<uint32_t *>(smem)[idx] = (uint32_t)(w & 0xffff);
    }
    __syncthreads();

    uint32_t go_s = O::splat(p.go), neg_ge = O::splat(-p.ge);
    uint32_t lane_off = (uint32_t)lane * 16u;
    uint32_t nz = lane != 0 ? 1u : 0u;
    uint32_t sym_stride_r = sym_stride;  // opaque: symbol index -> table offset by an IMAD (FMA pipe), not a shift (ALU)
    asm volatile("" : "+r"(go_s), "+r"(neg_ge), "+r"(lane_off), "+r"(nz), "+r"(sym_stride_r));
    const uint32_t pad_sym = (uint32_t)p.S;

    // a "task" = two streamed sequences (2t, 2t+1); a warp sweeps two tasks (chains a and b) per trip
    const uint32_t n_tasks = (p.n_rseq + 1) / 2;
    const uint32_t n_trips_total = (n_tasks + 1) / 2;
    const uint32_t total_warps = gridDim.x * warps_per_block;
    for (uint32_t trip = blockIdx.x * warps_per_block + warp; trip < n_trips_total; trip += total_warps) {
        uint32_t ids[4];
        int lens[4];
        uint64_t offs[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const uint64_t a = (uint64_t)4 * trip + h;
            uint32_t id = 0xffffffffu;
            if (a < p.n_rseq) id = p.task_ids ? p.task_ids[a] : (uint32_t)a;
            ids[h] = id;
            offs[h] = id != 0xffffffffu ? p.roff[id] : 0;
            lens[h] = id != 0xffffffffu ? (int)(p.roff[id + 1] - offs[h]) : 0;
        }
        const int LA = max(lens[0], lens[1]), LB = max(lens[2], lens[3]);

        // ---- stage the column symbols of the four sequences ----
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int Lh = h < 2 ? LA : LB;
            for (int j = lane; j < Lh; j += 32) {
                uint32_t sym = pad_sym;
                if (j < lens[h]) sym = (uint32_t)p.lut[p.rseq[offs[h] + j]];
                stage[h * stage_len + j] = (uint16_t)sym;
            }
        }
        __syncwarp();
        const uint16_t *cA_lo = stage, *cA_hi = stage + stage_len, *cB_lo = stage + 2 * stage_len, *cB_hi = stage + 3 * stage_len;

        uint32_t aH[2][K], aF[K], bH[2][K], bF[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            aH[0][i] = aH[1][i] = 0;
            aF[i] = 0;
            bH[0][i] = bH[1][i] = 0;
            bF[i] = 0;
        }
        uint32_t abest = 0, ah_last = 0, ae_out = 0, ah_up_prev = 0;
        uint32_t bbest = 0, bh_last = 0, be_out = 0, bh_up_prev = 0;
        const int nsteps = max(LA, LB) + (G - 1);

        // one column of one chain; STEADY = no activity test
        auto chain_col = [&](auto parity, auto which, const int j, const uint32_t h_in, const uint32_t e_in) {
            constexpr int PO = decltype(parity)::value, PN = 1 - PO;
            constexpr bool IS_A = decltype(which)::value == 0;
            const uint32_t off_lo = (uint32_t)(IS_A ? cA_lo[j] : cB_lo[j]) * sym_stride_r;
            const uint32_t off_hi = (uint32_t)(IS_A ? cA_hi[j] : cB_hi[j]) * sym_stride_r;
            const uint4 *tl = reinterpret_cast<const uint4 *>(smem + lane_off + off_lo);
            const uint4 *th = reinterpret_cast<const uint4 *>(smem + lane_off + off_hi);
            if (IS_A) {
                uint32_t adiag = ah_up_prev, aE = e_in, ahp = 0;
#pragma unroll
                for (int i4 = 0; i4 < K4; ++i4) {
                    const uint4 l4 = tl[i4 * G], h4 = th[i4 * G];
                    const uint32_t w[4] = {h4.x * 65536u + l4.x, h4.y * 65536u + l4.y, h4.z * 65536u + l4.z, h4.w * 65536u + l4.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = i4 * 4 + q;
                        if (i < K) ZOE_SCORE_ROW(a, w[q])
                    }
                }
                if (K & 1) abest = O::max2(abest, ahp);
                ah_last = aH[PN][K - 1];
                ae_out = aE;
            } else {
                uint32_t bdiag = bh_up_prev, bE = e_in, bhp = 0;
#pragma unroll
                for (int i4 = 0; i4 < K4; ++i4) {
                    const uint4 l4 = tl[i4 * G], h4 = th[i4 * G];
                    const uint32_t w[4] = {h4.x * 65536u + l4.x, h4.y * 65536u + l4.y, h4.z * 65536u + l4.z, h4.w * 65536u + l4.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = i4 * 4 + q;
                        if (i < K) ZOE_SCORE_ROW(b, w[q])
                    }
                }
                if (K & 1) bbest = O::max2(bbest, bhp);
                bh_last = bH[PN][K - 1];
                be_out = bE;
            }
            (void)h_in;
        };
        using C0 = std::integral_constant<int, 0>;
        using C1 = std::integral_constant<int, 1>;
        auto step_fn = [&](auto parity, auto steady, const int step) {
            const uint32_t ah_in = __shfl_up_sync(FULL, ah_last, 1) * nz, ae_in = __shfl_up_sync(FULL, ae_out, 1) * nz;
            const uint32_t bh_in = __shfl_up_sync(FULL, bh_last, 1) * nz, be_in = __shfl_up_sync(FULL, be_out, 1) * nz;
            const int j = step - lig;
            if (decltype(steady)::value || (j >= 0 && j < LA)) chain_col(parity, C0{}, j, ah_in, ae_in);
            if (decltype(steady)::value || (j >= 0 && j < LB)) chain_col(parity, C1{}, j, bh_in, be_in);
            ah_up_prev = ah_in;
            bh_up_prev = bh_in;
        };

        // segments: ramp (generic), steady while every lane of both chains has a column, tail (generic)
        const int Lmin = min(LA, LB);
        int step = 0;
        int seg_end[3] = {min(G, nsteps), Lmin - 1, nsteps};
#pragma unroll 1
        for (int seg = 0; seg < 3; ++seg) {
            const int end = seg_end[seg];
            if (seg == 1) {
                for (; step < end; step += 2) {
                    step_fn(C0{}, std::true_type{}, step);
                    step_fn(C1{}, std::true_type{}, step + 1);
                }
            } else {
                for (; step < end; ++step) {
                    if (step & 1)
                        step_fn(C1{}, std::false_type{}, step);
                    else
                        step_fn(C0{}, std::false_type{}, step);
                }
            }
        }

        // ---- reduce and write ----
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            abest = O::max2(abest, __shfl_xor_sync(FULL, abest, d));
            bbest = O::max2(bbest, __shfl_xor_sync(FULL, bbest, d));
        }
        if (lane == 0) {
            const int v[4] = {(int)(int16_t)(abest & 0xffff), (int)(int16_t)(abest >> 16), (int)(int16_t)(bbest & 0xffff),
                              (int)(int16_t)(bbest >> 16)};
#pragma unroll
            for (int h = 0; h < 4; ++h)
                if (ids[h] != 0xffffffffu) p.best[(size_t)ids[h] * p.n_cseq + rp.cj] = (v[h] >= p.ovf_thresh) ? -1 : v[h];
        }
    }
}

}  // namespace zoe_cuda
