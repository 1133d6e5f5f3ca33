// sw_align_win.cuh -- alignment with traceback in two passes over a checkpointed DP (sm_100a).
//
// Same contract as sw_align.cuh (zoe: sw_simd_align, src/alignment/sw/striped.rs:449-598 +
// BackTrackable::to_alignment, src/alignment/types/backtrack.rs:290-342), different cost model.  A traceback
// only ever consults cells between the alignment's start and end columns, so writing direction bits for the
// whole n x m matrix (what zoe's CPU kernel must do, and what sw_align_fill_kernel does) wastes > 85 % of the
// work when a 150-nt read is aligned against a 1.7-kb segment.  Here:
//
//   pass A  sw_align_scan_kernel     the score recurrence (4.5 ALU instr per packed cell pair, no direction
//                                    bits) over all columns.  Per lane it keeps, branch-free in packed 16-bit
//                                    halves, the running maximum and the first / last column pair in which the
//                                    maximum was reached, and it parks the systolic state every CB columns.
//   pairing win_classify / win_bucket_scan / win_scatter: a counting sort of the mapped pairs by
//                                    (profiled sequence, checkpoint block), so two pairs that restart from
//                                    the same column share one packed s16x2 task.
//   pass B  sw_align_winfill_kernel  restarts from the checkpoint at or before c_end - (r_end + 1 + slack),
//                                    recomputes that window with the five direction bits per cell, and pins the
//                                    best cell down to zoe's (max H, min r, min c) -- striped.rs:555-583 -- inside
//                                    the one column pair pass A pointed at.  The DP values inside the window are
//                                    bit-identical to the full matrix because the restart state is the full
//                                    matrix's own systolic state.
//   walk    sw_traceback_win_kernel  zoe's priority walk over the window; a walk that reaches the left edge
//                                    of its window (a gap longer than the slack) is handed to the literal
//                                    kernel, like a tie hazard.
//
// Which cell is "the" best cell?  zoe takes the smallest row, then the smallest column among the cells holding the
// maximum.  Lanes own disjoint ascending row ranges, so the winner lies in the lowest lane that reaches the
// maximum; inside that lane pass A knows the first and the last column pair holding it.  If they differ (typical
// for unrelated pairs, whose small maximum recurs) a smaller row could hide in a later column: a PIN sweep
// (sw_align_winfill_kernel<.., FLAGS=false> in pin mode) restarts at the checkpoint before the first occurrence,
// runs the plain recurrence up to the last one and finds the cell exactly; the pair then joins the others.
//
// Checkpoint layout per pass-A task and profiled sequence: ckpt[kb-1][w][lane] = one 32-bit word of the lane's
// vector (H[0..K), F[0..K), E leaving the last row, diagonal for the next column); kb = 1 .. (L-1)/CB
// is the systolic state before step kb*CB, where lane l is about to compute column kb*CB - l (CB >= G).
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "sw_align.cuh"

namespace zoe_cuda {

__host__ __device__ inline int ckpt_words_per_lane(int K) { return 2 * K + 2; }

struct WinParams {
    ScoreParams s;               // sequences, tables, scoring (s.best unused; s.n_rseq = sequences in the chunk)
    AlignEnd *ends;              // [n_rseq_total * n_cseq]
    uint32_t chunk_first;        // first batch sequence of the chunk
    // checkpoints (pass A writes, pass B reads)
    uint32_t *ckpt;
    const uint64_t *ckpt_base;   // [n_cseq] word offset of each profiled sequence inside a task's region
    uint64_t ckpt_task_stride;   // words per pass-A task
    int cb_log2;                 // checkpoint spacing CB = 1 << cb_log2 columns
    uint32_t slack;              // extra columns kept left of the shortest possible walk
    // pairing
    uint32_t nblk;               // checkpoint blocks per profiled sequence (key = cj * nblk + block)
    uint32_t *hist;              // [n_cseq * nblk] bucket sizes, then bucket cursors
    uint32_t *bucket_start;      // [n_cseq * nblk + 1] even-aligned exclusive scan
    uint32_t *items;             // sorted global pair ids, 0xffffffff = empty slot
    uint32_t *n_items;           // [1] total slots (even)
    // pass B
    uint32_t *flags;             // window flag words
    uint64_t win_task_stride;    // words per pass-B task = Wmax * G * NW
    uint32_t wmax;               // window capacity in columns
    unsigned long long *counters; // [7] internal consistency failures
    // reverse instantiation of pass A (ranges, sw_ranges.cuh): tasks = pairs of `items` sharing (profiled, end column)
    int rev_maxw;                // largest substitution weight (bounds how many columns a reversed sub-problem needs)
    const AlignEnd *rev_in;      // forward end cells (exact)
    AlignEnd *rev_out;           // (best, r', c') of the reversed, truncated sub-problem
    int pin_mode;                // 0: window fill; 1: pin sweep, result back in pass A's representation;
                                 // 2: pin sweep, exact (row, column) published (ranges)
};

// Pass A leaves (winning lane, even step of the first column pair); the best cell is one of that lane's K rows in
// columns (s - lane, s + 1 - lane).  Upper bounds of its row / column, used to size the window before pass B
// pins the cell down.
__device__ __forceinline__ void end_bounds(uint32_t lane, uint32_t s_even, int K, uint32_t n, uint32_t L, uint32_t &r_hi,
                                           uint32_t &c_hi) {
    r_hi = min(lane * (uint32_t)K + (uint32_t)K - 1u, n - 1u);
    c_hi = min(s_even + 1u - lane, L - 1u);  // the lane was active in the pair, so s_even + 1 >= lane
}

__device__ __forceinline__ uint32_t win_start(uint32_t r_end, uint32_t c_end, uint32_t slack, int cb_log2) {
    const uint32_t need = r_end + 1 + slack;  // columns the walk may touch: #M <= r_end + 1, #I + 1 <= slack
    const uint32_t lo = (c_end + 1 > need) ? (c_end + 1 - need) : 0;
    return (lo >> cb_log2) << cb_log2;
}

// ---------------------------------------------------------------------------------------------
// pass A: scores, end cells, checkpoints
// ---------------------------------------------------------------------------------------------
// Same register-resident systolic sweep as sw_score_kernel (ping-pong H sets, no register moves).  Three things
// keep the per-step overhead off the ALU pipe, which bounds the kernel:
//   * steps G-1 .. L-1 have every lane active, so the steady loop carries no per-lane activity predicates;
//   * the best-cell bookkeeping (max H, then min r, then min c -- striped.rs:555-583) is tested once per two
//     columns with a packed "any half >= running best" test and handled in a cold branch;
//   * checkpoints are taken at a warp-uniform step boundary (the whole systolic state, skewed by lane), so the
//     test is a uniform-datapath compare.
// State parked before step s0 = kb*CB (s0 even, s0 >= G): lane l holds column s0-1-l in H[0], its along-row gap
// state F, the E leaving its last row and the diagonal it received at step s0-1.
// CSM = the profiled symbol codes are staged in shared memory (LDS with a 32-bit address in the hot loop); false = a
// profiled set too large for that (> 96 KB): the codes are read from global memory (L1-resident: every warp of the
// SM walks the same sequences).
// REV = the reverse pass of sw_simd_score_ranges (striped.rs:355-388): a task is two (streamed, profiled) items that
// share the profiled sequence and the end column; rows = each item's reversed prefix streamed[r_end], streamed[r_end-1],
// ..., columns = profiled[c_end], profiled[c_end-1], ...; no checkpoints -- the reversed alignment starts in the
// corner, so the exact best cell is pinned by re-sweeping the first columns from scratch inside the same kernel.
#ifndef ZOE_SCAN_THREADS
#define ZOE_SCAN_THREADS 640
#endif
template <int G, int K, bool CSM = true, bool REV = false>
__global__ void __launch_bounds__(K > 24 ? 256 : (K > 16 && K <= 20 && !REV ? ZOE_SCAN_THREADS : 512)) sw_align_scan_kernel(const WinParams wp) {
    using O = Ops<true>;
    const ScoreParams &p = wp.s;
    constexpr int K4 = (K + 3) / 4;
    constexpr int CKW = 2 * K + 2;  // checkpoint words per lane
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) uint8_t smem[];

    const int tid = threadIdx.x;
    const int lig = tid % G;
    const int group_in_block = tid / G;
    const int groups_per_block = blockDim.x / G;

    const TaskSmem sm = carve_and_stage<G, K4>(smem, p, CSM);
    const int tab_bytes = sm.tab_bytes;
    const uint8_t *cc = CSM ? sm.s_cc : p.ccodes;

    uint32_t go_s = O::splat(p.go), neg_ge = O::splat(-p.ge);
    const uint32_t cb_mask = (1u << wp.cb_log2) - 1u;
    const bool lane0 = lig == 0;  // lane 0 receives zeros from "above"
    uint32_t nz_lane = lane0 ? 0u : 1u;
    asm volatile("" : "+r"(nz_lane));
    const uint32_t one_s = O::splat(1);
    // loop invariants the compiler would otherwise rematerialise inside the hot loop (S2R + address arithmetic);
    // the table offset stays a 32-bit offset into the shared array so the loads remain LDS.128
    uint32_t tab_lane_off = (uint32_t)group_in_block * (uint32_t)tab_bytes + (uint32_t)lig * 16u;
    asm volatile("" : "+r"(go_s), "+r"(neg_ge), "+r"(tab_lane_off));

    const uint32_t n_tasks = REV ? *wp.n_items / 2 : p.n_tasks;
    const uint32_t total_groups = gridDim.x * groups_per_block;
    const uint32_t trips = (n_tasks + total_groups - 1) / total_groups;
    const uint32_t first = blockIdx.x * groups_per_block + group_in_block;

    for (uint32_t trip = 0; trip < trips; ++trip) {
        const uint32_t task = first + trip * total_groups;
        const bool valid = task < n_tasks;
        uint32_t id_lo = 0xffffffffu, id_hi = 0xffffffffu;
        uint32_t g_lo = 0xffffffffu, g_hi = 0xffffffffu;  // REV: global pair ids
        uint64_t off_lo = 0, off_hi = 0;                   // byte offset of row 0
        int len_lo = 0, len_hi = 0;
        uint32_t rev_cj = 0;
        int rev_cend = -1, rev_cols = 0;
        // Columns a reversed sub-problem can need.  Every cell of the reversed matrix that holds the forward score S is
        // the start of an optimal alignment ending in the forward end cell (that cell is the first one holding S, so
        // nothing else inside the prefix rectangle reaches S).  Such an alignment consumes n_r <= r_end + 1 rows and
        // n_c = (diagonal moves) + I columns with  S <= n_r * maxw - go - (I - 1) * ge  for I >= 1 (one gap run is the
        // cheapest way to spend I column-only moves), hence  n_c <= n_r + 1 + (n_r * maxw - go - S) / ge.  Columns
        // beyond that cannot hold S, so the sweep stops there: zoe's reverse pass (striped.rs:355-388) sweeps the whole
        // prefix and finds the same first cell.  ge == 0: no bound.
        auto rev_bound = [&](const AlignEnd &e) -> int {
            if (p.ge <= 0) return 0x7fffffff;
            const long long nr = (long long)e.r_end + 1;
            const long long num = nr * (long long)wp.rev_maxw - (long long)p.go - (long long)e.best;
            const long long ic = num >= 0 ? 1 + num / (long long)p.ge : 0;
            const long long w = nr + ic + 1;
            return w > 0x7fffffffLL ? 0x7fffffff : (int)w;
        };
        if (!REV) {
            if (valid) {
                const uint32_t a = 2 * task, b = 2 * task + 1;
                id_lo = wp.chunk_first + a;
                id_hi = (b < p.n_rseq) ? wp.chunk_first + b : 0xffffffffu;
            }
            if (id_lo != 0xffffffffu) {
                off_lo = p.roff[id_lo];
                len_lo = (int)(p.roff[id_lo + 1] - off_lo);
            }
            if (id_hi != 0xffffffffu) {
                off_hi = p.roff[id_hi];
                len_hi = (int)(p.roff[id_hi + 1] - off_hi);
            }
        } else {
            if (valid) {
                g_lo = wp.items[2 * task];
                g_hi = wp.items[2 * task + 1];
            }
            if (g_lo != 0xffffffffu) {
                const AlignEnd e = wp.rev_in[g_lo];
                id_lo = g_lo / p.n_cseq;
                rev_cj = g_lo % p.n_cseq;
                rev_cend = (int)e.c_end;
                off_lo = p.roff[id_lo] + e.r_end;  // row r' reads streamed[r_end - r']
                len_lo = (int)e.r_end + 1;
                rev_cols = rev_bound(e);
            }
            if (g_hi != 0xffffffffu) {
                const AlignEnd e = wp.rev_in[g_hi];
                id_hi = g_hi / p.n_cseq;
                off_hi = p.roff[id_hi] + e.r_end;
                len_hi = (int)e.r_end + 1;
                rev_cols = max(rev_cols, rev_bound(e));
            }
        }

        build_task_table<G, K, true>(sm, p, lig, (int64_t)off_lo, len_lo, (int64_t)off_hi, len_hi, REV ? -1 : 1);

        uint32_t *ck_task = REV ? nullptr : wp.ckpt + (size_t)task * wp.ckpt_task_stride;

        for (uint32_t sweep = 0; sweep < (REV ? 1u : p.n_cseq); ++sweep) {
            const uint32_t cj = REV ? rev_cj : sweep;
            const uint32_t c0 = p.coff[cj];
            // REV: the groups of a warp sweep different numbers of columns (sorted by end column, so nearly equal)
            const int L = REV ? min(rev_cend + 1, rev_cols) : (int)(p.coff[cj + 1] - c0);
            const uint8_t *cs = cc + c0 + (REV ? max(rev_cend, 0) : 0);  // column j reads cs[j] (REV: cs[-j])
            // the lane's column of step s is s - lig: fold the lane into the pointer once per sweep (the compiler would
            // otherwise re-derive lig from %tid.x inside the steady loop, in front of the column-code load)
            const uint8_t *cs_lane = REV ? cs + lig : cs - lig;
            uint32_t cs_idx = CSM ? (uint32_t)(cs_lane - smem) : 0u;  // staged columns: an index into the shared array (LDS)
            if (CSM)
                asm volatile("" : "+r"(cs_idx));
            else
                asm volatile("" : "+l"(cs_lane));
            uint32_t *ck = REV ? nullptr : ck_task + wp.ckpt_base[cj] + lig;

            uint32_t H[2][K], F[K];
#pragma unroll
            for (int i = 0; i < K; ++i) {
                H[0][i] = H[1][i] = 0;
                F[i] = 0;
            }
            uint32_t h_last = 0, e_out = 0, h_up_prev = 0;
            // branch-free best bookkeeping, all packed per 16-bit half: running maximum, and the first / last
            // step pair (s >> 1) in which this lane's column maximum reached it
            uint32_t bestp = 0, firstp = 0, lastp = 0;
            uint32_t cm = 0;               // maximum of the column(s) since the last bookkeeping
            uint32_t hp_pend = 0;          // odd K, steady loop: last row of an even step, folded in by the next step
            const int nsteps = REV ? __reduce_max_sync(FULL, L > 0 ? L + G - 1 : 0) : L + G - 1;
            const int L_steady = REV ? __reduce_min_sync(FULL, L) : L;  // every lane of every group has a column below it

            // one column of this lane: K rows, H read from set PO (column j-1), written to set PN
            // STEADY (the two-step steady loop only): with an odd K the column's last row is not folded into the column
            // maximum by an extra VIMNMX; it stays pending in `hp_pend` and pairs up with row 0 of the next (odd) step,
            // so two steps fold their 2K values with exactly K VIMNMX3.
            // `st` is the STEP (the lane is folded into cs_lane / cs_idx)
            auto column = [&](auto parity, auto steady, const int st, const uint32_t h_in, const uint32_t e_in) {
                constexpr int PO = decltype(parity)::value, PN = 1 - PO;
                constexpr bool STEADY = decltype(steady)::value;
                constexpr int PH = (STEADY && (K & 1)) ? PO : 0;  // pairing phase of this column
                const uint32_t code = CSM ? smem[cs_idx + (uint32_t)(REV ? -st : st)] : cs_lane[REV ? -st : st];
                const uint4 *tp = reinterpret_cast<const uint4 *>(smem + tab_lane_off + code * (uint32_t)(K4 * G * 16));
                uint32_t diag = h_up_prev, E = e_in, hp = hp_pend;
#pragma unroll
                for (int i4 = 0; i4 < K4; ++i4) {
                    const uint4 w4 = tp[i4 * G];
                    const uint32_t wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = i4 * 4 + q;
                        if (i < K) {
                            const uint32_t x = O::max3(E, F[i], go_s) - go_s;
                            const uint32_t Hn = O::addmax(diag, wv[q], x);
                            diag = H[PO][i];
                            E = O::addmax(E, neg_ge, Hn);
                            F[i] = O::addmax(F[i], neg_ge, Hn);
                            H[PN][i] = Hn;
                            if ((i + PH) & 1)
                                cm = O::max3(cm, Hn, hp);
                            else
                                hp = Hn;
                        }
                    }
                }
                if (K & 1) {
                    if (!STEADY)
                        cm = O::max2(cm, hp);
                    else if (PH == 0)
                        hp_pend = hp;
                }
                h_last = H[PN][K - 1];
                e_out = E;
            };
            using SteadyT = std::true_type;
            using GenericT = std::false_type;
            using I0 = std::integral_constant<int, 0>;
            using I1 = std::integral_constant<int, 1>;
            // m - bestp and m - cm are >= 0 in both halves, so the plain 32-bit subtractions are exact (FMA pipe);
            // the 0/1 halves are widened to 0xffff masks by a multiply (FMA pipe); 5 ALU instructions in all.
            auto bookkeeping = [&](const uint32_t s_even) {
                const uint32_t pp = (s_even >> 1) * 0x00010001u;
                const uint32_t m = O::max2(cm, bestp);
                const uint32_t inc = O::min2(m - bestp, one_s) * 0xffffu;            // halves that grew
                const uint32_t ge = (one_s - O::min2(m - cm, one_s)) * 0xffffu;      // halves with cm >= bestp
                firstp = (firstp & ~inc) | (pp & inc);
                lastp = (lastp & ~ge) | (pp & ge);
                bestp = m;
                cm = 0;
            };
            auto generic_step = [&](auto parity, const int s) {
                const uint32_t h_sh = __shfl_up_sync(FULL, h_last, 1, G), h_in = lane0 ? 0u : h_sh;
                const uint32_t e_sh = __shfl_up_sync(FULL, e_out, 1, G), e_in = lane0 ? 0u : e_sh;
                const int j = s - lig;
                if (j >= 0 && j < L) {
                    column(parity, GenericT{}, s, h_in, e_in);
                    bookkeeping((uint32_t)s & ~1u);
                }
                h_up_prev = h_in;
            };
            auto checkpoint = [&](const int s) {  // the state before step s (s even: the last column is in set 0)
                // 32-bit stores on purpose: a 128-bit store would pin H/F to aligned register quads in the hot loop
                uint32_t *dst = ck + (size_t)(((uint32_t)s >> wp.cb_log2) - 1) * (CKW * G);
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    dst[i * G] = H[0][i];
                    dst[(K + i) * G] = F[i];
                }
                dst[(2 * K) * G] = e_out;
                dst[(2 * K + 1) * G] = h_up_prev;
            };

            int s = 0;
            const int s_steady = min(G, nsteps);  // G is even and >= the first step with every lane active
            for (; s < s_steady; ++s) {
                if (s & 1)
                    generic_step(I1{}, s);
                else
                    generic_step(I0{}, s);
            }
            while (s + 1 < L_steady) {  // steps s and s+1: every lane has a column
                if (!REV && (((uint32_t)s) & cb_mask) == 0 && valid) checkpoint(s);
                // the inner loop runs to the next checkpoint boundary without a branch in its body (CB >= G >= 4, even)
                const int s_end = REV ? L_steady : min(L_steady, (int)(((uint32_t)s | cb_mask) + 1u));
                for (; s + 1 < s_end; s += 2) {
                    {   // lane 0 receives zeros from "above": a multiply on the FMA pipe, not a SEL on the ALU pipe
                        const uint32_t h_in = __shfl_up_sync(FULL, h_last, 1, G) * nz_lane;
                        const uint32_t e_in = __shfl_up_sync(FULL, e_out, 1, G) * nz_lane;
                        column(I0{}, SteadyT{}, s, h_in, e_in);
                        h_up_prev = h_in;
                    }
                    {
                        const uint32_t h_in = __shfl_up_sync(FULL, h_last, 1, G) * nz_lane;
                        const uint32_t e_in = __shfl_up_sync(FULL, e_out, 1, G) * nz_lane;
                        column(I1{}, SteadyT{}, s + 1, h_in, e_in);
                        h_up_prev = h_in;
                    }
                    bookkeeping((uint32_t)s);
                }
                if (s_end >= L_steady) break;
            }
            for (; s < nsteps; ++s) {
                if (!REV && (((uint32_t)s) & cb_mask) == 0 && s >= G && s < L && valid) checkpoint(s);
                if (s & 1)
                    generic_step(I1{}, s);
                else
                    generic_step(I0{}, s);
            }

            // ---- the lowest lane holding the group's maximum wins (smallest rows); fetch its first / last pair ----
            uint32_t key_lo = ((bestp & 0xffffu) << 8) | (uint32_t)(G - 1 - lig);
            uint32_t key_hi = ((bestp >> 16) << 8) | (uint32_t)(G - 1 - lig);
#pragma unroll
            for (int d = G / 2; d >= 1; d >>= 1) {
                key_lo = max(key_lo, __shfl_xor_sync(FULL, key_lo, d, G));
                key_hi = max(key_hi, __shfl_xor_sync(FULL, key_hi, d, G));
            }
            const int wl_lo = G - 1 - (int)(key_lo & 0xffu), wl_hi = G - 1 - (int)(key_hi & 0xffu);
            const uint32_t f_lo = __shfl_sync(FULL, firstp, wl_lo, G) & 0xffffu, l_lo = __shfl_sync(FULL, lastp, wl_lo, G) & 0xffffu;
            const uint32_t f_hi = __shfl_sync(FULL, firstp, wl_hi, G) >> 16, l_hi = __shfl_sync(FULL, lastp, wl_hi, G) >> 16;
            if (REV) {
                // ---- pin: re-sweep the first columns from scratch; the winning lane looks for the maximum (smallest
                //      row, then smallest column) in the steps between its first and last occurrence ----
                const int b_lo = (int)(key_lo >> 8), b_hi = (int)(key_hi >> 8);
                const int fs_lo = 2 * (int)f_lo, se_lo = 2 * (int)l_lo + 1, fs_hi = 2 * (int)f_hi, se_hi = 2 * (int)l_hi + 1;
                int last_step = -1;
                if (g_lo != 0xffffffffu && b_lo > 0) last_step = se_lo;
                if (g_hi != 0xffffffffu && b_hi > 0) last_step = max(last_step, se_hi);
                const int n_re = __reduce_max_sync(FULL, last_step + 1);
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    H[0][i] = H[1][i] = 0;
                    F[i] = 0;
                }
                h_last = e_out = h_up_prev = 0;
                int bi_lo = K, bi_hi = K, bj_lo = 0, bj_hi = 0;
                auto pin_step = [&](auto parity, const int st) {
                    constexpr int PN = 1 - decltype(parity)::value;
                    const uint32_t h_sh = __shfl_up_sync(FULL, h_last, 1, G), h_in = lane0 ? 0u : h_sh;
                    const uint32_t e_sh = __shfl_up_sync(FULL, e_out, 1, G), e_in = lane0 ? 0u : e_sh;
                    const int j = st - lig;
                    const bool in_cols = j >= 0 && j < L;
                    if (in_cols) column(parity, GenericT{}, st, h_in, e_in);
                    // the search for the best cell sits behind a warp-uniform branch (see sw_align_winfill_kernel)
                    const bool hit_lo = in_cols && lig == wl_lo && st >= fs_lo && st <= se_lo;
                    const bool hit_hi = in_cols && lig == wl_hi && st >= fs_hi && st <= se_hi;
                    if (__any_sync(FULL, hit_lo || hit_hi)) {
                        if (hit_lo) {
                            int irow = K;
#pragma unroll
                            for (int i = K - 1; i >= 0; --i)
                                if ((int)(int16_t)(H[PN][i] & 0xffffu) == b_lo) irow = i;
                            if (irow < bi_lo) {
                                bi_lo = irow;
                                bj_lo = j;
                            }
                        }
                        if (hit_hi) {
                            int irow = K;
#pragma unroll
                            for (int i = K - 1; i >= 0; --i)
                                if ((int)(int16_t)(H[PN][i] >> 16) == b_hi) irow = i;
                            if (irow < bi_hi) {
                                bi_hi = irow;
                                bj_hi = j;
                            }
                        }
                    }
                    h_up_prev = h_in;
                };
                for (int st = 0; st < n_re; ++st) {
                    if (st & 1)
                        pin_step(I1{}, st);
                    else
                        pin_step(I0{}, st);
                }
                if (g_lo != 0xffffffffu && lig == wl_lo) {
                    AlignEnd e;
                    e.best = b_lo;
                    e.r_end = (uint32_t)(wl_lo * K + bi_lo);
                    e.c_end = (uint32_t)bj_lo;
                    e.aux = 0;
                    if (b_lo > 0 && bi_lo >= K) atomicAdd(&wp.counters[7], 1ULL);  // scan and pin disagree: internal error
                    wp.rev_out[g_lo] = e;
                }
                if (g_hi != 0xffffffffu && lig == wl_hi) {
                    AlignEnd e;
                    e.best = b_hi;
                    e.r_end = (uint32_t)(wl_hi * K + bi_hi);
                    e.c_end = (uint32_t)bj_hi;
                    e.aux = 0;
                    if (b_hi > 0 && bi_hi >= K) atomicAdd(&wp.counters[7], 1ULL);
                    wp.rev_out[g_hi] = e;
                }
            } else if (lig == 0 && valid) {
                if (id_lo != 0xffffffffu) {
                    AlignEnd e;
                    const int b = (int)(key_lo >> 8);
                    e.best = (b >= p.ovf_thresh) ? -1 : b;
                    e.r_end = (uint32_t)wl_lo;   // winning lane, refined to a row by pass B
                    e.c_end = 2u * f_lo;         // even step of the pair: columns (s - lane, s + 1 - lane)
                    e.aux = 2u * l_lo;           // last pair holding the maximum in that lane (== c_end: unambiguous)
                    wp.ends[(size_t)id_lo * p.n_cseq + cj] = e;
                }
                if (id_hi != 0xffffffffu) {
                    AlignEnd e;
                    const int b = (int)(key_hi >> 8);
                    e.best = (b >= p.ovf_thresh) ? -1 : b;
                    e.r_end = (uint32_t)wl_hi;
                    e.c_end = 2u * f_hi;
                    e.aux = 2u * l_hi;
                    wp.ends[(size_t)id_hi * p.n_cseq + cj] = e;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pass A, two tasks per group: the forward scan above with TWO independent recurrences per thread.
// ---------------------------------------------------------------------------------------------
// With one profiled sequence per sweep (config 3) sw_align_scan_kernel carries one dependency chain per thread and
// stalls on it (ncu: `wait` 1.9 cycles per issue, ALU pipe 83 %).  Here a group sweeps tasks 2t and 2t+1 side by side:
// same column symbol, two score tables, two H / F register sets -- the instruction-level parallelism the two-stream score
// kernel gets from its second profiled sequence.  Twice the shared memory per group, so 11 warps per SM instead of 16,
// but 22 chains instead of 16.  Checkpoints, bookkeeping and the published (lane, step pair) are exactly those of the
// one-task kernel, task by task, so the pin sweep, pass B and the walk do not change.
template <int K>
struct ScanChain {
    uint32_t H[2][K], F[K];
    uint32_t h_last, e_out, h_up_prev;
    uint32_t bestp, firstp, lastp, cm, hp_pend;
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            H[0][i] = H[1][i] = 0;
            F[i] = 0;
        }
        h_last = e_out = h_up_prev = 0;
        bestp = firstp = lastp = cm = hp_pend = 0;
    }
};

#define ZOE_SCAN_ROW(C, W)                                            \
    {                                                                 \
        const uint32_t x = O::max3(C##E, C.F[i], go_s) - go_s;        \
        const uint32_t Hn = O::addmax(C##diag, W, x);                 \
        C##diag = C.H[PO][i];                                         \
        C##E = O::addmax(C##E, neg_ge, Hn);                           \
        C.F[i] = O::addmax(C.F[i], neg_ge, Hn);                       \
        C.H[PN][i] = Hn;                                              \
        if ((i + PH) & 1)                                             \
            C.cm = O::max3(C.cm, Hn, C##hp);                          \
        else                                                          \
            C##hp = Hn;                                               \
    }

template <int G, int K, bool CSM = true>
__global__ void __launch_bounds__(K <= 13 ? 512 : 352) sw_align_scan2_kernel(const WinParams wp) {
    using O = Ops<true>;
    const ScoreParams &p = wp.s;
    constexpr int K4 = (K + 3) / 4;
    constexpr int CKW = 2 * K + 2;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) uint8_t smem[];

    const int tid = threadIdx.x;
    const int lig = tid % G;
    const int group_in_block = tid / G;
    const int groups_per_block = blockDim.x / G;

    const TaskSmem sm0 = carve_and_stage<G, K4, 2>(smem, p, CSM);
    const int tab_bytes = sm0.tab_bytes;
    TaskSmem smA = sm0, smB = sm0;
    smB.tab = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(sm0.tab) + tab_bytes);
    const uint8_t *cc = CSM ? sm0.s_cc : p.ccodes;

    uint32_t go_s = O::splat(p.go), neg_ge = O::splat(-p.ge);
    const uint32_t cb_mask = (1u << wp.cb_log2) - 1u;
    const bool lane0 = lig == 0;
    uint32_t nz_lane = lane0 ? 0u : 1u;
    const uint32_t one_s = O::splat(1);
    uint32_t tabA_off = (uint32_t)group_in_block * (uint32_t)(2 * tab_bytes) + (uint32_t)lig * 16u;
    uint32_t tabB_off = tabA_off + (uint32_t)tab_bytes;
    asm volatile("" : "+r"(go_s), "+r"(neg_ge), "+r"(tabA_off), "+r"(tabB_off), "+r"(nz_lane));

    const uint32_t n_duos = (p.n_tasks + 1) / 2;
    const uint32_t total_groups = gridDim.x * groups_per_block;
    const uint32_t trips = (n_duos + total_groups - 1) / total_groups;
    const uint32_t first = blockIdx.x * groups_per_block + group_in_block;

    for (uint32_t trip = 0; trip < trips; ++trip) {
        const uint32_t duo = first + trip * total_groups;
        uint32_t task[2] = {2 * duo, 2 * duo + 1};
        bool valid[2];
        uint32_t id_lo[2], id_hi[2];
        uint64_t off_lo[2] = {0, 0}, off_hi[2] = {0, 0};
        int len_lo[2] = {0, 0}, len_hi[2] = {0, 0};
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            valid[t] = duo < n_duos && task[t] < p.n_tasks;
            id_lo[t] = id_hi[t] = 0xffffffffu;
            if (valid[t]) {
                const uint32_t a = 2 * task[t], b = 2 * task[t] + 1;
                id_lo[t] = wp.chunk_first + a;
                id_hi[t] = (b < p.n_rseq) ? wp.chunk_first + b : 0xffffffffu;
            }
            if (id_lo[t] != 0xffffffffu) {
                off_lo[t] = p.roff[id_lo[t]];
                len_lo[t] = (int)(p.roff[id_lo[t] + 1] - off_lo[t]);
            }
            if (id_hi[t] != 0xffffffffu) {
                off_hi[t] = p.roff[id_hi[t]];
                len_hi[t] = (int)(p.roff[id_hi[t] + 1] - off_hi[t]);
            }
        }
        build_task_table<G, K, true>(smA, p, lig, (int64_t)off_lo[0], len_lo[0], (int64_t)off_hi[0], len_hi[0]);
        build_task_table<G, K, true>(smB, p, lig, (int64_t)off_lo[1], len_lo[1], (int64_t)off_hi[1], len_hi[1]);
        uint32_t *ckA_task = wp.ckpt + (size_t)task[0] * wp.ckpt_task_stride;
        uint32_t *ckB_task = wp.ckpt + (size_t)task[1] * wp.ckpt_task_stride;

        for (uint32_t cj = 0; cj < p.n_cseq; ++cj) {
            const uint32_t c0 = p.coff[cj];
            const int L = (int)(p.coff[cj + 1] - c0);
            const uint8_t *cs = cc + c0;
            uint32_t *ckA = ckA_task + wp.ckpt_base[cj] + lig;
            uint32_t *ckB = ckB_task + wp.ckpt_base[cj] + lig;

            ScanChain<K> A, B;
            A.reset();
            B.reset();
            const int nsteps = L + G - 1;

            auto column2 = [&](auto parity, auto steady, const int j, const uint32_t ha_in, const uint32_t ea_in,
                               const uint32_t hb_in, const uint32_t eb_in) {
                constexpr int PO = decltype(parity)::value, PN = 1 - PO;
                constexpr bool STEADY = decltype(steady)::value;
                constexpr int PH = (STEADY && (K & 1)) ? PO : 0;
                const uint32_t code_off = (uint32_t)cs[j] * (uint32_t)(K4 * G * 16);
                const uint4 *tpA = reinterpret_cast<const uint4 *>(smem + tabA_off + code_off);
                const uint4 *tpB = reinterpret_cast<const uint4 *>(smem + tabB_off + code_off);
                uint32_t Adiag = A.h_up_prev, AE = ea_in, Ahp = A.hp_pend;
                uint32_t Bdiag = B.h_up_prev, BE = eb_in, Bhp = B.hp_pend;
#pragma unroll
                for (int i4 = 0; i4 < K4; ++i4) {
                    const uint4 wa4 = tpA[i4 * G], wb4 = tpB[i4 * G];
                    const uint32_t wa[4] = {wa4.x, wa4.y, wa4.z, wa4.w};
                    const uint32_t wb[4] = {wb4.x, wb4.y, wb4.z, wb4.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = i4 * 4 + q;
                        if (i < K) {
                            ZOE_SCAN_ROW(A, wa[q])
                            ZOE_SCAN_ROW(B, wb[q])
                        }
                    }
                }
                if (K & 1) {
                    if (!STEADY) {
                        A.cm = O::max2(A.cm, Ahp);
                        B.cm = O::max2(B.cm, Bhp);
                    } else if (PH == 0) {
                        A.hp_pend = Ahp;
                        B.hp_pend = Bhp;
                    }
                }
                A.h_last = A.H[PN][K - 1];
                A.e_out = AE;
                B.h_last = B.H[PN][K - 1];
                B.e_out = BE;
            };
            using SteadyT = std::true_type;
            using GenericT = std::false_type;
            using I0 = std::integral_constant<int, 0>;
            using I1 = std::integral_constant<int, 1>;
            auto bookkeeping = [&](ScanChain<K> &C, const uint32_t s_even) {
                const uint32_t pp = (s_even >> 1) * 0x00010001u;
                const uint32_t m = O::max2(C.cm, C.bestp);
                const uint32_t inc = O::min2(m - C.bestp, one_s) * 0xffffu;
                const uint32_t ge = (one_s - O::min2(m - C.cm, one_s)) * 0xffffu;
                C.firstp = (C.firstp & ~inc) | (pp & inc);
                C.lastp = (C.lastp & ~ge) | (pp & ge);
                C.bestp = m;
                C.cm = 0;
            };
            auto generic_step = [&](auto parity, const int s) {
                const uint32_t ha = __shfl_up_sync(FULL, A.h_last, 1, G) * nz_lane, ea = __shfl_up_sync(FULL, A.e_out, 1, G) * nz_lane;
                const uint32_t hb = __shfl_up_sync(FULL, B.h_last, 1, G) * nz_lane, eb = __shfl_up_sync(FULL, B.e_out, 1, G) * nz_lane;
                const int j = s - lig;
                if (j >= 0 && j < L) {
                    column2(parity, GenericT{}, j, ha, ea, hb, eb);
                    bookkeeping(A, (uint32_t)s & ~1u);
                    bookkeeping(B, (uint32_t)s & ~1u);
                }
                A.h_up_prev = ha;
                B.h_up_prev = hb;
            };
            auto checkpoint = [&](const ScanChain<K> &C, uint32_t *ck, const int s) {
                uint32_t *dst = ck + (size_t)(((uint32_t)s >> wp.cb_log2) - 1) * (CKW * G);
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    dst[i * G] = C.H[0][i];
                    dst[(K + i) * G] = C.F[i];
                }
                dst[(2 * K) * G] = C.e_out;
                dst[(2 * K + 1) * G] = C.h_up_prev;
            };

            int s = 0;
            const int s_steady = min(G, nsteps);
            for (; s < s_steady; ++s) {
                if (s & 1)
                    generic_step(I1{}, s);
                else
                    generic_step(I0{}, s);
            }
            for (; s + 1 < L; s += 2) {
                if ((((uint32_t)s) & cb_mask) == 0) {
                    if (valid[0]) checkpoint(A, ckA, s);
                    if (valid[1]) checkpoint(B, ckB, s);
                }
                {
                    const uint32_t ha = __shfl_up_sync(FULL, A.h_last, 1, G) * nz_lane, ea = __shfl_up_sync(FULL, A.e_out, 1, G) * nz_lane;
                    const uint32_t hb = __shfl_up_sync(FULL, B.h_last, 1, G) * nz_lane, eb = __shfl_up_sync(FULL, B.e_out, 1, G) * nz_lane;
                    column2(I0{}, SteadyT{}, s - lig, ha, ea, hb, eb);
                    A.h_up_prev = ha;
                    B.h_up_prev = hb;
                }
                {
                    const uint32_t ha = __shfl_up_sync(FULL, A.h_last, 1, G) * nz_lane, ea = __shfl_up_sync(FULL, A.e_out, 1, G) * nz_lane;
                    const uint32_t hb = __shfl_up_sync(FULL, B.h_last, 1, G) * nz_lane, eb = __shfl_up_sync(FULL, B.e_out, 1, G) * nz_lane;
                    column2(I1{}, SteadyT{}, s + 1 - lig, ha, ea, hb, eb);
                    A.h_up_prev = ha;
                    B.h_up_prev = hb;
                }
                bookkeeping(A, (uint32_t)s);
                bookkeeping(B, (uint32_t)s);
            }
            for (; s < nsteps; ++s) {
                if ((((uint32_t)s) & cb_mask) == 0 && s >= G && s < L) {
                    if (valid[0]) checkpoint(A, ckA, s);
                    if (valid[1]) checkpoint(B, ckB, s);
                }
                if (s & 1)
                    generic_step(I1{}, s);
                else
                    generic_step(I0{}, s);
            }

            // ---- per task: the lowest lane holding the group's maximum wins; publish (lane, first / last step pair) ----
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const uint32_t bestp = t == 0 ? A.bestp : B.bestp, firstp = t == 0 ? A.firstp : B.firstp,
                               lastp = t == 0 ? A.lastp : B.lastp;
                uint32_t key_lo = ((bestp & 0xffffu) << 8) | (uint32_t)(G - 1 - lig);
                uint32_t key_hi = ((bestp >> 16) << 8) | (uint32_t)(G - 1 - lig);
#pragma unroll
                for (int d = G / 2; d >= 1; d >>= 1) {
                    key_lo = max(key_lo, __shfl_xor_sync(FULL, key_lo, d, G));
                    key_hi = max(key_hi, __shfl_xor_sync(FULL, key_hi, d, G));
                }
                const int wl_lo = G - 1 - (int)(key_lo & 0xffu), wl_hi = G - 1 - (int)(key_hi & 0xffu);
                const uint32_t f_lo = __shfl_sync(FULL, firstp, wl_lo, G) & 0xffffu, l_lo = __shfl_sync(FULL, lastp, wl_lo, G) & 0xffffu;
                const uint32_t f_hi = __shfl_sync(FULL, firstp, wl_hi, G) >> 16, l_hi = __shfl_sync(FULL, lastp, wl_hi, G) >> 16;
                if (lig == 0 && valid[t]) {
                    if (id_lo[t] != 0xffffffffu) {
                        AlignEnd e;
                        const int b = (int)(key_lo >> 8);
                        e.best = (b >= p.ovf_thresh) ? -1 : b;
                        e.r_end = (uint32_t)wl_lo;
                        e.c_end = 2u * f_lo;
                        e.aux = 2u * l_lo;
                        wp.ends[(size_t)id_lo[t] * p.n_cseq + cj] = e;
                    }
                    if (id_hi[t] != 0xffffffffu) {
                        AlignEnd e;
                        const int b = (int)(key_hi >> 8);
                        e.best = (b >= p.ovf_thresh) ? -1 : b;
                        e.r_end = (uint32_t)wl_hi;
                        e.c_end = 2u * f_hi;
                        e.aux = 2u * l_hi;
                        wp.ends[(size_t)id_hi[t] * p.n_cseq + cj] = e;
                    }
                }
            }
        }
    }
}
#undef ZOE_SCAN_ROW

// ---------------------------------------------------------------------------------------------
// pairing: classify every pair of the chunk, counting-sort the mapped ones by (profiled, block)
// ---------------------------------------------------------------------------------------------
struct ClassifyParams {
    AlignEnd *ends;
    const uint64_t *roff;
    const uint32_t *coff;
    uint32_t n_cseq, chunk_first, n_slots;
    int K;
    int cb_log2;
    uint32_t slack, nblk;
    int maxw, go, ge;              // largest weight and the (positive) gap penalties: bound the walk's column-only moves
    int all_exact;
    uint32_t *hist;
    uint8_t *amb;                  // [chunk-local pair] 1 = the pair's end cell is ambiguous (written by pin stage 1)
    int subset;                    // classification proper: 0 = every pair, 1 = unambiguous pairs only, 2 = ambiguous only
                                   // (the pin sweep of the ambiguous ones overlaps the window fill of the others)
    int pin_stage;                 // 1: select the ambiguous pairs for the pin sweep; 2: every mapped pair (ranges);
                                   // 0: the classification proper
    int32_t *best_arr;
    uint32_t *score;
    uint8_t *status, *tier, *hazard;
    uint32_t *ref_start, *ref_end, *query_start, *query_end, *cig_count;
    unsigned long long *counters;  // [5] exact-list length, [8] packed overflows, [11] pinned (ambiguous) pairs
    uint32_t *hazard_list;
    TierPolicy tp;
};

// Restart step of a pin sweep: the checkpoint at or before the first occurrence (0 = from scratch).  Pass A parks
// its state only at steps < L (a step index can exceed L-1 by the lane skew), hence the clamp.
__device__ __forceinline__ uint32_t pin_start(uint32_t first_s, uint32_t L, int cb_log2) {
    return (min(first_s, L - 1u) >> cb_log2) << cb_log2;
}

// Does this pair take part in the windowed pipeline at all (mapped, inside the packed range, inside the widest
// allowed integer type, not literal-only)?
__device__ __forceinline__ bool win_eligible(const AlignEnd &e, uint32_t n, const TierPolicy &tp, int all_exact) {
    return e.best > 0 && n != 0 && !all_exact && tier_for(tp, (uint32_t)e.best) != 0;
}

__global__ void win_classify_kernel(const ClassifyParams t) {
    const uint32_t pairs = t.n_slots * t.n_cseq;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pairs) return;
    const uint32_t seq_local = k / t.n_cseq, cj = k % t.n_cseq;
    const uint32_t seq = t.chunk_first + seq_local;
    const size_t gid = (size_t)seq * t.n_cseq + cj;
    if (!t.pin_stage && t.subset && (t.amb[k] != 0) != (t.subset == 2)) return;  // the other pass's pair: not even read
    const AlignEnd e = t.ends[gid];
    const uint32_t n = (uint32_t)(t.roff[seq + 1] - t.roff[seq]);
    if (t.pin_stage) {  // ambiguous end cells only: bucket them by (profiled sequence, checkpoint block)
        const bool sel = win_eligible(e, n, t.tp, t.all_exact) && (t.pin_stage == 2 || e.aux != e.c_end);
        if (t.amb) t.amb[k] = sel ? 1 : 0;
        if (sel) {
            atomicAdd(&t.hist[cj * t.nblk + (pin_start(e.c_end, t.coff[cj + 1] - t.coff[cj], t.cb_log2) >> t.cb_log2)], 1u);
            if (e.aux != e.c_end) atomicAdd(&t.counters[11], 1ULL);
        }
        return;
    }
    t.hazard[gid] = 0;
    t.cig_count[gid] = 0;
    t.best_arr[gid] = e.best;
    if (e.best < 0) {  // packed lanes reached the overflow threshold: exact 32-bit score + literal kernel
        t.status[gid] = 0xFF;
        t.hazard[gid] = 1;
        atomicAdd(&t.counters[8], 1ULL);
        unsigned long long slot = atomicAdd(&t.counters[5], 1ULL);
        t.hazard_list[slot] = (uint32_t)gid | 0x80000000u;  // no exact end cell yet
        t.ends[gid].aux = kNotBucketed;
        return;
    }
    if (e.best == 0 || n == 0) {
        t.score[gid] = 0;
        t.status[gid] = 2;  // Unmapped
        t.tier[gid] = t.tp.first;
        t.ref_start[gid] = t.ref_end[gid] = t.query_start[gid] = t.query_end[gid] = 0;
        t.ends[gid].aux = kNotBucketed;
        return;
    }
    const uint8_t tier = tier_for(t.tp, (uint32_t)e.best);
    if (tier == 0) {  // beyond the widest allowed integer type: Overflowed
        t.score[gid] = 0;
        t.status[gid] = 1;
        t.tier[gid] = t.tp.last;
        t.ref_start[gid] = t.ref_end[gid] = t.query_start[gid] = t.query_end[gid] = 0;
        t.ends[gid].aux = kNotBucketed;
        return;
    }
    t.score[gid] = (uint32_t)e.best;
    t.status[gid] = 0;
    t.tier[gid] = tier;
    if (t.all_exact) {
        unsigned long long slot = atomicAdd(&t.counters[5], 1ULL);
        t.hazard_list[slot] = (uint32_t)gid | 0x80000000u;  // pass A's (lane, step pair) is not an end cell
        t.hazard[gid] = 1;
        t.ends[gid].aux = kNotBucketed;
        return;
    }
    uint32_t r_hi, c_hi;
    end_bounds(e.r_end, e.c_end, t.K, n, t.coff[cj + 1] - t.coff[cj], r_hi, c_hi);
    // The walk follows one alignment of score S over n_r <= r_hi + 1 rows; its column-only moves I satisfy
    // S <= n_r * maxw - go - (I - 1) * ge (one gap run is the cheapest way to spend them), so I + 1 columns of slack are
    // enough whenever that is less than the configured slack (a mapped 150-nt read: 2 instead of 16).  Were the bound
    // ever too tight the walk would reach the window's left edge and the pair would go to the literal kernel.
    uint32_t slack = t.slack;
    if (t.ge > 0) {
        const long long num = (long long)(r_hi + 1) * t.maxw - t.go - (long long)e.best;
        const long long imax = num >= 0 ? 1 + num / t.ge : 0;
        if (imax + 1 < (long long)slack) slack = (uint32_t)(imax + 1);
    }
    const uint32_t ws = win_start(r_hi, c_hi, slack, t.cb_log2);
    t.ends[gid].aux = ws;
    atomicAdd(&t.hist[cj * t.nblk + (ws >> t.cb_log2)], 1u);
}

// Single-block exclusive scan of the bucket sizes rounded up to even (a pass-B task holds two pairs of ONE
// bucket); leaves hist zeroed so the scatter kernel can use it as the per-bucket cursor.
__global__ void win_bucket_scan_kernel(uint32_t *hist, uint32_t n_keys, uint32_t *bucket_start, uint32_t *n_items) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_keys; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        uint32_t v = 0;
        if (i < n_keys) {
            v = (hist[i] + 1u) & ~1u;
            hist[i] = 0;
        }
        uint32_t incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0u;
            uint32_t wi = w;
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t excl = s_carry + s_warp[wid] + incl - v;
        if (i < n_keys) bucket_start[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        bucket_start[n_keys] = s_carry;
        *n_items = s_carry;
    }
}

__global__ void win_scatter_kernel(const ClassifyParams t, const uint32_t *bucket_start, uint32_t *items) {
    const uint32_t pairs = t.n_slots * t.n_cseq;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pairs) return;
    const uint32_t seq = t.chunk_first + k / t.n_cseq, cj = k % t.n_cseq;
    const size_t gid = (size_t)seq * t.n_cseq + cj;
    if (!t.pin_stage && t.subset && (t.amb[k] != 0) != (t.subset == 2)) return;
    const AlignEnd e = t.ends[gid];
    uint32_t key;
    if (t.pin_stage) {
        const uint32_t n = (uint32_t)(t.roff[seq + 1] - t.roff[seq]);
        if (!(win_eligible(e, n, t.tp, t.all_exact) && (t.pin_stage == 2 || e.aux != e.c_end))) return;
        key = cj * t.nblk + (pin_start(e.c_end, t.coff[cj + 1] - t.coff[cj], t.cb_log2) >> t.cb_log2);
    } else {
        if (e.aux == kNotBucketed) return;  // unmapped / overflowed / literal-only
        key = cj * t.nblk + (e.aux >> t.cb_log2);
    }
    const uint32_t slot = bucket_start[key] + atomicAdd(&t.hist[key], 1u);
    items[slot] = (uint32_t)gid;
}

// ---------------------------------------------------------------------------------------------
// pass B: direction bits for the window [ws, c_end] of every mapped pair            (FLAGS = true,  wp.pin_mode = 0)
// pin   : exact best cell of the pairs whose maximum recurs in the winning lane      (FLAGS = false, wp.pin_mode = 1)
// Both resume the systolic sweep from a checkpoint of pass A and search the winning lane for the best cell (smallest
// row, then smallest column -- striped.rs:555-583) inside a range of steps; only pass B emits direction bits.
// ---------------------------------------------------------------------------------------------
template <int G, int K, bool FLAGS = true>
__global__ void __launch_bounds__(512) sw_align_winfill_kernel(const WinParams wp) {
    using O = Ops<true>;
    const ScoreParams &p = wp.s;
    constexpr int K4 = (K + 3) / 4;
    constexpr int CKW = 2 * K + 2;
    constexpr int NW = (((K + 2) / 3) + 3) & ~3;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) uint8_t smem[];

    const int tid = threadIdx.x;
    const int lig = tid % G;
    const int group_in_block = tid / G;
    const int groups_per_block = blockDim.x / G;

    const TaskSmem sm = carve_and_stage<G, K4>(smem, p, p.cols_in_smem != 0);
    uint4 *const tab = sm.tab;
    const uint8_t *cc = p.cols_in_smem ? sm.s_cc : p.ccodes;

    const uint32_t go_s = O::splat(p.go), neg_ge = O::splat(-p.ge), one_s = O::splat(1), c21 = O::splat(21);

    const uint32_t n_tasks = *wp.n_items / 2;
    const uint32_t total_groups = gridDim.x * groups_per_block;
    const uint32_t trips = (n_tasks + total_groups - 1) / total_groups;
    const uint32_t first = blockIdx.x * groups_per_block + group_in_block;

    for (uint32_t trip = 0; trip < trips; ++trip) {
        const uint32_t task = first + trip * total_groups;
        const bool valid = task < n_tasks;
        uint32_t g_lo = 0xffffffffu, g_hi = 0xffffffffu;
        if (valid) {
            g_lo = wp.items[2 * task];
            g_hi = wp.items[2 * task + 1];
        }
        // the two pairs share (profiled sequence, restart step); a bucket of odd size leaves g_hi empty
        uint32_t cj = 0, ws = 0;
        int we = -1;       // last column computed
        int s_last = -1;   // last step of the sweep (absolute)
        uint32_t se_lo = 0, se_hi = 0;  // last step searched for the best cell (the first one is fs_*)
        uint64_t off_lo = 0, off_hi = 0;
        int len_lo = 0, len_hi = 0;
        uint32_t seqA_lo = 0, seqA_hi = 0;  // chunk-relative sequence index (locates the checkpoint)
        // where pass A saw the maximum: (winning lane, even step of the column pair), and its value
        uint32_t ls_lo = 0xffffffffu, ls_hi = 0xffffffffu, fs_lo = 0, fs_hi = 0;
        int best_lo = -1, best_hi = -1;
        if (g_lo != 0xffffffffu) {
            const uint32_t seq = g_lo / p.n_cseq;
            cj = g_lo % p.n_cseq;
            const AlignEnd e = wp.ends[g_lo];
            off_lo = p.roff[seq];
            len_lo = (int)(p.roff[seq + 1] - off_lo);
            const uint32_t Lc = p.coff[cj + 1] - p.coff[cj];
            if (wp.pin_mode) {  // from the checkpoint before the first occurrence to the last one, all columns
                ws = pin_start(e.c_end, Lc, wp.cb_log2);
                we = (int)Lc - 1;
                se_lo = e.aux + 1;
                s_last = (int)se_lo;
            } else {
                uint32_t r_hi, c_hi;
                end_bounds(e.r_end, e.c_end, K, (uint32_t)len_lo, Lc, r_hi, c_hi);
                ws = e.aux;
                we = (int)c_hi;
                se_lo = e.c_end + 1;
                s_last = we + G - 1;
            }
            ls_lo = e.r_end;
            fs_lo = e.c_end;
            best_lo = e.best;
            seqA_lo = seq - wp.chunk_first;
        }
        if (g_hi != 0xffffffffu) {
            const uint32_t seq = g_hi / p.n_cseq;
            const AlignEnd e = wp.ends[g_hi];
            off_hi = p.roff[seq];
            len_hi = (int)(p.roff[seq + 1] - off_hi);
            const uint32_t Lc = p.coff[cj + 1] - p.coff[cj];
            if (wp.pin_mode) {
                se_hi = e.aux + 1;
                s_last = max(s_last, (int)se_hi);
            } else {
                uint32_t r_hi, c_hi;
                end_bounds(e.r_end, e.c_end, K, (uint32_t)len_hi, Lc, r_hi, c_hi);
                we = max(we, (int)c_hi);
                se_hi = e.c_end + 1;
                s_last = max(s_last, (int)c_hi + G - 1);
            }
            ls_hi = e.r_end;
            fs_hi = e.c_end;
            best_hi = e.best;
            seqA_hi = seq - wp.chunk_first;
        }
        int bi_lo = K, bi_hi = K, bj_lo = 0, bj_hi = 0;  // best cell inside the pointed-at column pair

        build_task_table<G, K, true>(sm, p, lig, (int64_t)off_lo, len_lo, (int64_t)off_hi, len_hi);

        // ---- restore the systolic state before step ws (sw_align_scan_kernel's checkpoint layout) ----
        uint32_t Hrow[K], Frow[K];
        uint32_t e_out = 0, h_up_prev = 0;
        {
            uint32_t v[CKW];
#pragma unroll
            for (int i = 0; i < CKW; ++i) v[i] = 0;
            if (ws > 0) {
                const uint32_t kb = ws >> wp.cb_log2;
                if (g_lo != 0xffffffffu) {
                    const uint32_t *src = wp.ckpt + (size_t)(seqA_lo >> 1) * wp.ckpt_task_stride + wp.ckpt_base[cj] +
                                          (size_t)(kb - 1) * (CKW * G) + lig;
                    const int sh = (seqA_lo & 1) ? 16 : 0;
#pragma unroll
                    for (int i = 0; i < CKW; ++i) v[i] |= (src[i * G] >> sh) & 0xffffu;
                }
                if (g_hi != 0xffffffffu) {
                    const uint32_t *src = wp.ckpt + (size_t)(seqA_hi >> 1) * wp.ckpt_task_stride + wp.ckpt_base[cj] +
                                          (size_t)(kb - 1) * (CKW * G) + lig;
                    const int sh = (seqA_hi & 1) ? 16 : 0;
#pragma unroll
                    for (int i = 0; i < CKW; ++i) v[i] |= ((src[i * G] >> sh) & 0xffffu) << 16;
                }
            }
#pragma unroll
            for (int i = 0; i < K; ++i) {
                Hrow[i] = v[i];
                Frow[i] = v[K + i];
            }
            e_out = v[2 * K];
            h_up_prev = v[2 * K + 1];
        }

        const uint8_t *cs = cc + p.coff[cj];
        uint32_t *fl = wp.flags + (size_t)task * wp.win_task_stride + (size_t)lig * NW;
        // The sweep resumes at step ws: lane l continues with column ws - l, so the flag buffer starts at column
        // ws - (G-1); only columns >= ws are complete (and only those are consulted by the walk).
        const int origin = (int)ws - (G - 1);
        uint32_t h_last = Hrow[K - 1];
        // the groups of a warp hold different windows: the trip count must be warp-uniform for the shuffles
        const int nsteps = __reduce_max_sync(FULL, s_last >= (int)ws ? s_last - (int)ws + 1 : 0);

        for (int step = 0; step < nsteps; ++step) {
            uint32_t h_in = __shfl_up_sync(FULL, h_last, 1, G);
            uint32_t e_in = __shfl_up_sync(FULL, e_out, 1, G);
            if (lig == 0) {
                h_in = 0;
                e_in = 0;
            }
            const int j = (int)ws + step - lig;  // column of this lane
            if (j >= 0 && j <= we) {
                const int jw = j - origin;
                const uint4 *tp = tab + (size_t)cs[j] * (K4 * G) + lig;
                uint32_t diag = h_up_prev;
                uint32_t E = e_in;
                uint32_t words[FLAGS ? NW : 1];
                if (FLAGS) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) words[w] = 0;
                }
#pragma unroll
                for (int i4 = 0; i4 < K4; ++i4) {
                    const uint4 w4 = tp[i4 * G];
                    const uint32_t wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = i4 * 4 + q;
                        if (i < K) {
                            const uint32_t Fi = Frow[i];
                            uint32_t x = O::max3(E, Fi, go_s) - go_s;
                            uint32_t H = O::addmax(diag, wv[q], x);
                            diag = Hrow[i];
                            uint32_t E2 = O::addmax(E, neg_ge, H);
                            uint32_t F2 = O::addmax(Fi, neg_ge, H);
                            if (FLAGS) {
                                uint32_t Hg = H + go_s;
                                uint32_t nve = O::min2(Hg - E, one_s);    // 0 where E == H   (UP)
                                uint32_t nhe = O::min2(Hg - Fi, one_s);   // 0 where F == H   (LEFT)
                                uint32_t xv = O::min2(E2 - H, one_s);     // 1 where next-row E extends (UP_EXT)
                                uint32_t xh = O::min2(F2 - H, one_s);     // 1 where next-col F extends (LEFT_EXT)
                                uint32_t ns = O::min2(H, one_s);          // 0 where H == 0   (STOP)
                                uint32_t code = c21 + 2u * xv + 8u * xh - nve - 4u * nhe - 16u * ns;
                                words[i / 3] += code << (5 * (i % 3));
                            }
                            E = E2;
                            Frow[i] = F2;
                            Hrow[i] = H;
                        }
                    }
                }
                h_last = Hrow[K - 1];
                e_out = E;
                if (FLAGS && valid) {
#pragma unroll
                    for (int w = 0; w < NW; w += 4)
                        *reinterpret_cast<uint4 *>(fl + (size_t)jw * (G * NW) + w) =
                            make_uint4(words[w], words[w + 1], words[w + 2], words[w + 3]);
                }
            }
            // ---- pin the best cell down: smallest row, then smallest column (striped.rs:555-583).  One lane of a group
            //      looks, in the two steps of its column pair (pin mode: between the first and last occurrence), so the
            //      search sits behind a warp-uniform branch: as straight-line predicated code it was 114 of the ~290
            //      ALU instructions of EVERY step (ncu: 15 ALU instructions per cell pair instead of 9.5) ----
            {
                const uint32_t s_abs = ws + (uint32_t)step;
                const bool in_win = j >= 0 && j <= we;
                const bool hit_lo = in_win && (uint32_t)lig == ls_lo && s_abs >= fs_lo && s_abs <= se_lo;
                const bool hit_hi = in_win && (uint32_t)lig == ls_hi && s_abs >= fs_hi && s_abs <= se_hi;
                if (__any_sync(FULL, hit_lo || hit_hi)) {
                    if (hit_lo) {
                        int irow = K;
#pragma unroll
                        for (int i = K - 1; i >= 0; --i)
                            if ((int)(int16_t)(Hrow[i] & 0xffffu) == best_lo) irow = i;
                        if (irow < bi_lo) {
                            bi_lo = irow;
                            bj_lo = j;
                        }
                    }
                    if (hit_hi) {
                        int irow = K;
#pragma unroll
                        for (int i = K - 1; i >= 0; --i)
                            if ((int)(int16_t)(Hrow[i] >> 16) == best_hi) irow = i;
                        if (irow < bi_hi) {
                            bi_hi = irow;
                            bj_hi = j;
                        }
                    }
                }
            }
            h_up_prev = h_in;
        }
        // the winning lane publishes the exact end cell (the walk starts there)
        // (pin mode: the cell goes back into pass A's representation -- winning lane, even step of the column
        //  pair -- now unambiguous, so the pair joins the regular classification)
        if ((uint32_t)lig == ls_lo) {
            if (bi_lo < K) {
                if (wp.pin_mode == 1) {
                    wp.ends[g_lo].c_end = ((uint32_t)bj_lo + ls_lo) & ~1u;
                    wp.ends[g_lo].aux = ((uint32_t)bj_lo + ls_lo) & ~1u;
                } else {
                    wp.ends[g_lo].r_end = ls_lo * K + (uint32_t)bi_lo;
                    wp.ends[g_lo].c_end = (uint32_t)bj_lo;
                }
            } else {
                atomicAdd(&wp.counters[7], 1ULL);  // pass A and pass B disagree: reported as an internal error
            }
        }
        if ((uint32_t)lig == ls_hi) {
            if (bi_hi < K) {
                if (wp.pin_mode == 1) {
                    wp.ends[g_hi].c_end = ((uint32_t)bj_hi + ls_hi) & ~1u;
                    wp.ends[g_hi].aux = ((uint32_t)bj_hi + ls_hi) & ~1u;
                } else {
                    wp.ends[g_hi].r_end = ls_hi * K + (uint32_t)bi_hi;
                    wp.ends[g_hi].c_end = (uint32_t)bj_hi;
                }
            } else {
                atomicAdd(&wp.counters[7], 1ULL);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// walk over the window flags
// ---------------------------------------------------------------------------------------------
struct TraceWinParams {
    TraceParams t;               // outputs, lengths, hazard list (t.flags / t.flag_base / t.task_stride unused)
    const uint32_t *items;
    const uint32_t *n_items;
    const uint32_t *wflags;
    uint64_t win_task_stride;
    int cb_log2;
    uint32_t slack;
};

__global__ void sw_traceback_win_kernel(const TraceWinParams tw) {
    const TraceParams &t = tw.t;
    const uint32_t pslot = blockIdx.x * blockDim.x + threadIdx.x;
    if (pslot >= *tw.n_items) return;
    const uint32_t gid = tw.items[pslot];
    if (gid == 0xffffffffu) return;
    const uint32_t seq = gid / t.n_cseq, cj = gid % t.n_cseq;
    const AlignEnd e = t.ends[gid];
    const uint32_t n = (uint32_t)(t.roff[seq + 1] - t.roff[seq]);
    const uint32_t m = t.coff[cj + 1] - t.coff[cj];
    const uint32_t ws = e.aux;
    const uint32_t half = pslot & 1u;
    const uint32_t *fl = tw.wflags + (size_t)(pslot >> 1) * tw.win_task_stride;
    const int G = t.G, K = t.K, NW = t.NW;
    auto cell = [&](uint32_t r, uint32_t c) -> uint32_t {
        uint32_t lane = r / K, kk = r % K;
        uint32_t w = fl[((size_t)(c - ws + (uint32_t)(G - 1)) * G + lane) * NW + kk / 3];
        return (w >> (16 * half + 5 * (kk % 3))) & 31u;
    };

    const uint32_t k = gid - t.chunk_first * t.n_cseq;  // chunk-local pair index (CIGAR scratch slot)
    CigarBack cg;
    cg.init(t.cig_scratch + (size_t)k * t.cig_cap, t.cig_cap);
    const uint32_t OP_UP = t.invert ? 1u /*I*/ : 2u /*D*/, OP_LEFT = t.invert ? 2u : 1u;

    uint32_t r = e.r_end + 1, c = e.c_end + 1;
    const uint32_t r_end1 = r, c_end1 = c;
    cg.push(4u, t.invert ? (n - r_end1) : (m - c_end1));
    uint32_t cur = cell(e.r_end, e.c_end);
    int op = 0;  // 0 none, 1 D(up), 2 I(left), 3 M
    bool hz = false;
    while (!(cur & 16u) && r > 0 && c > 0) {
        if ((cur & 1u) && (cur & 4u)) hz = true;
        if (op == 1 && (cur & 2u)) {
            r -= 1;
        } else if (op == 2 && (cur & 8u)) {
            c -= 1;
        } else if (cur & 1u) {
            op = 1;
            r -= 1;
        } else if (cur & 4u) {
            op = 2;
            c -= 1;
        } else {
            op = 3;
            r -= 1;
            c -= 1;
        }
        cg.push(op == 1 ? OP_UP : (op == 2 ? OP_LEFT : 0u), 1);
        if (r == 0 || c == 0) break;       // zoe reads cell(max(r-1,0), max(c-1,0)) and then leaves the loop
        if (c - 1 < ws) {                  // the walk left its window: hand the pair to the literal kernel
            hz = true;
            atomicAdd(&t.counters[10], 1ULL);
            break;
        }
        cur = cell(r - 1, c - 1);
    }
    cg.push(4u, t.invert ? r : c);  // leading soft clip
    cg.flush();

    if (hz) {
        unsigned long long slot = atomicAdd(&t.counters[5], 1ULL);
        t.hazard_list[slot] = gid;
        t.hazard[gid] = 1;
        return;  // the exact kernel rewrites everything for this pair
    }
    if (cg.overflow) atomicAdd(&t.counters[6], 1ULL);
    t.cig_count[gid] = cg.n;
    if (t.invert) {  // ref_range <-> query_range (output.rs:418-419)
        t.ref_start[gid] = c;
        t.ref_end[gid] = c_end1;
        t.query_start[gid] = r;
        t.query_end[gid] = r_end1;
    } else {
        t.ref_start[gid] = r;
        t.ref_end[gid] = r_end1;
        t.query_start[gid] = c;
        t.query_end[gid] = c_end1;
    }
}

}  // namespace zoe_cuda
