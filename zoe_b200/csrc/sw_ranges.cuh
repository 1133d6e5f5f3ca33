// sw_ranges.cuh -- score + alignment ranges without a traceback matrix (sm_100a).
//
// Replaces zoe's sw_simd_score_ends / sw_simd_score_ranges (src/alignment/sw/striped.rs:153-162, 213-336, 355-388) and
// the escalation chain ProfileSets::sw_score_ranges_from_i8 (src/alignment/profile_set.rs:313-359).
//
// zoe's recipe: a forward pass yields the score and the end cell (max H, then min row, then min column --
// striped.rs:296-321); a reverse pass over reference[..ref_end] and the reversed profile of query[..query_end]
// (StripedProfile::reverse_from_forward, src/alignment/profile.rs:314-350) yields the start cell the same way.
// Both passes are the score recurrence plus best-cell bookkeeping; no direction bits are stored, so memory stays
// O(pairs) and the function also serves pairs whose full matrix would not fit.
//
// Here both passes are one kernel, sw_ends_kernel<G,K,PACKED>:
//   forward   tasks = consecutive pairs of streamed sequences, every profiled sequence swept left to right;
//   reverse   tasks = pairs of (streamed, profiled) items that share the profiled sequence AND the end column
//             (counting sort by (profiled, c_end): win_bucket_scan_kernel / ranges_scatter_kernel), so the two
//             16-bit halves still see the same column symbols: profiled[c_end], profiled[c_end-1], ...; the rows
//             are each item's own reversed prefix streamed[r_end], streamed[r_end-1], ...
// End cells are independent of zoe's striped lane layout (they are properties of the H matrix), so no tie hazards
// arise here.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "sw_align.cuh"

namespace zoe_cuda {

struct EndsParams {
    ScoreParams s;             // sequences, tables, scoring (s.best unused; forward: s.n_rseq = sequences in the chunk)
    uint32_t chunk_first;      // forward: first batch sequence of the chunk
    const uint32_t *items;     // reverse: sorted global pair ids (0xffffffff = empty slot); nullptr = forward
    const uint32_t *n_items;   // reverse: number of slots (even)
    const AlignEnd *ends_in;   // reverse: the forward pass's end cells
    AlignEnd *out;             // forward: end cells; reverse: (best, r', c') of the reversed sub-problem
};

template <int G, int K, bool PACKED>
__global__ void __launch_bounds__(K > 24 ? 256 : 512) sw_ends_kernel(const EndsParams ep) {
    using O = Ops<PACKED>;
    const ScoreParams &p = ep.s;
    constexpr int K4 = (K + 3) / 4;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) uint8_t smem[];

    const int tid = threadIdx.x;
    const int lig = tid % G;
    const int group_in_block = tid / G;
    const int groups_per_block = blockDim.x / G;

    const TaskSmem sm = carve_and_stage<G, K4>(smem, p, p.cols_in_smem != 0);
    uint4 *const tab = sm.tab;
    const uint8_t *cc = p.cols_in_smem ? sm.s_cc : p.ccodes;

    const uint32_t go_s = O::splat(p.go), neg_ge = O::splat(-p.ge);
    const bool reverse = ep.items != nullptr;

    const uint32_t n_tasks = reverse ? (PACKED ? *ep.n_items / 2 : *ep.n_items) : p.n_tasks;
    const uint32_t total_groups = gridDim.x * groups_per_block;
    const uint32_t trips = (n_tasks + total_groups - 1) / total_groups;
    const uint32_t first = blockIdx.x * groups_per_block + group_in_block;

    for (uint32_t trip = 0; trip < trips; ++trip) {
        const uint32_t task = first + trip * total_groups;
        const bool valid = task < n_tasks;
        // ---- the task's two row sequences: base pointer, direction, length; and its column range ----
        uint32_t g_lo = 0xffffffffu, g_hi = 0xffffffffu;   // reverse: global pair ids
        uint32_t id_lo = 0xffffffffu, id_hi = 0xffffffffu; // streamed sequence ids
        int64_t base_lo = 0, base_hi = 0;                  // byte offset of row 0
        int dir = 1;                                       // +1 forward, -1 reverse (rows and columns alike)
        int len_lo = 0, len_hi = 0;
        uint32_t cj_first = 0, cj_last = 0;                // profiled sequences to sweep: [cj_first, cj_last)
        int rev_cend = 0;
        if (valid) {
            if (!reverse) {
                const uint32_t a = PACKED ? 2 * task : task, b = 2 * task + 1;
                id_lo = ep.chunk_first + a;
                if (PACKED && b < p.n_rseq) id_hi = ep.chunk_first + b;
                cj_last = p.n_cseq;
            } else {
                g_lo = ep.items[PACKED ? 2 * task : task];
                if (PACKED) g_hi = ep.items[2 * task + 1];
                dir = -1;
            }
        }
        if (!reverse) {
            if (id_lo != 0xffffffffu) {
                base_lo = (int64_t)p.roff[id_lo];
                len_lo = (int)(p.roff[id_lo + 1] - p.roff[id_lo]);
            }
            if (id_hi != 0xffffffffu) {
                base_hi = (int64_t)p.roff[id_hi];
                len_hi = (int)(p.roff[id_hi + 1] - p.roff[id_hi]);
            }
        } else {
            if (g_lo != 0xffffffffu) {
                const AlignEnd e = ep.ends_in[g_lo];
                id_lo = g_lo / p.n_cseq;
                cj_first = g_lo % p.n_cseq;
                cj_last = cj_first + 1;
                rev_cend = (int)e.c_end;
                base_lo = (int64_t)p.roff[id_lo] + e.r_end;  // row r' reads streamed[r_end - r']
                len_lo = (int)e.r_end + 1;
            }
            if (g_hi != 0xffffffffu) {
                const AlignEnd e = ep.ends_in[g_hi];
                id_hi = g_hi / p.n_cseq;
                base_hi = (int64_t)p.roff[id_hi] + e.r_end;
                len_hi = (int)e.r_end + 1;
            }
        }

        build_task_table<G, K, PACKED>(sm, p, lig, base_lo, len_lo, base_hi, len_hi, dir);

        // the groups of a warp may sweep different numbers of profiled sequences / columns: keep the loops uniform
        const uint32_t n_sweeps = __reduce_max_sync(FULL, cj_last - cj_first);
        for (uint32_t sw = 0; sw < n_sweeps; ++sw) {
            const uint32_t cj = cj_first + sw;
            const bool sweep_valid = cj < cj_last;
            int L = 0;
            const uint8_t *cs = cc;
            if (sweep_valid) {
                const uint32_t c0 = p.coff[cj];
                L = reverse ? rev_cend + 1 : (int)(p.coff[cj + 1] - c0);
                cs = cc + c0 + (reverse ? rev_cend : 0);  // column j reads cs[dir * j]
            }

            uint32_t Hrow[K], Frow[K];
#pragma unroll
            for (int i = 0; i < K; ++i) {
                Hrow[i] = 0;
                Frow[i] = 0;
            }
            uint32_t h_last = 0, e_out = 0, h_up_prev = 0;
            int bv_lo = 0, bv_hi = 0, bi_lo = 0, bi_hi = 0, bj_lo = 0, bj_hi = 0;
            const int nsteps = __reduce_max_sync(FULL, L > 0 ? L + G - 1 : 0);

            for (int step = 0; step < nsteps; ++step) {
                uint32_t h_in = __shfl_up_sync(FULL, h_last, 1, G);
                uint32_t e_in = __shfl_up_sync(FULL, e_out, 1, G);
                if (lig == 0) {
                    h_in = 0;
                    e_in = 0;
                }
                const int j = step - lig;
                if (j >= 0 && j < L) {
                    const int s = cs[dir * j];
                    const uint4 *tp = tab + (size_t)s * (K4 * G) + lig;
                    uint32_t diag = h_up_prev;
                    uint32_t E = e_in;
                    uint32_t cm = 0, hprev = 0;
#pragma unroll
                    for (int i4 = 0; i4 < K4; ++i4) {
                        const uint4 w4 = tp[i4 * G];
                        const uint32_t wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int i = i4 * 4 + q;
                            if (i < K) {
                                uint32_t x = O::max3(E, Frow[i], go_s) - go_s;
                                uint32_t H = O::addmax(diag, wv[q], x);
                                diag = Hrow[i];
                                E = O::addmax(E, neg_ge, H);
                                Frow[i] = O::addmax(Frow[i], neg_ge, H);
                                Hrow[i] = H;
                                if (i & 1)
                                    cm = O::max3(cm, H, hprev);
                                else
                                    hprev = H;
                            }
                        }
                    }
                    if (K & 1) cm = O::max2(cm, hprev);
                    h_last = Hrow[K - 1];
                    e_out = E;

                    // ---- best-cell bookkeeping: (max H, min r, min c), striped.rs:296-321 ----
                    const int cm_lo = PACKED ? (int)(int16_t)(cm & 0xffff) : (int)cm;
                    const int cm_hi = PACKED ? (int)(int16_t)(cm >> 16) : 0;
                    if (cm_lo > 0 && cm_lo >= bv_lo) {
                        int irow = K;
#pragma unroll
                        for (int i = K - 1; i >= 0; --i) {
                            int h = PACKED ? (int)(int16_t)(Hrow[i] & 0xffff) : (int)Hrow[i];
                            if (h == cm_lo) irow = i;
                        }
                        if (cm_lo > bv_lo || irow < bi_lo) {
                            bv_lo = cm_lo;
                            bi_lo = irow;
                            bj_lo = j;
                        }
                    }
                    if (PACKED && cm_hi > 0 && cm_hi >= bv_hi) {
                        int irow = K;
#pragma unroll
                        for (int i = K - 1; i >= 0; --i) {
                            int h = (int)(int16_t)(Hrow[i] >> 16);
                            if (h == cm_hi) irow = i;
                        }
                        if (cm_hi > bv_hi || irow < bi_hi) {
                            bv_hi = cm_hi;
                            bi_hi = irow;
                            bj_hi = j;
                        }
                    }
                }
                h_up_prev = h_in;
            }

            // ---- group reduction of (best, r, c): larger score, then smaller row, then smaller column ----
            unsigned long long key_lo = ((unsigned long long)(uint32_t)bv_lo << 40) |
                                        ((unsigned long long)(0xFFFFFu - (uint32_t)(lig * K + bi_lo)) << 20) |
                                        (unsigned long long)(0xFFFFFu - (uint32_t)bj_lo);
            unsigned long long key_hi = ((unsigned long long)(uint32_t)bv_hi << 40) |
                                        ((unsigned long long)(0xFFFFFu - (uint32_t)(lig * K + bi_hi)) << 20) |
                                        (unsigned long long)(0xFFFFFu - (uint32_t)bj_hi);
#pragma unroll
            for (int d = G / 2; d >= 1; d >>= 1) {
                unsigned long long o = __shfl_xor_sync(FULL, key_lo, d, G);
                key_lo = o > key_lo ? o : key_lo;
                o = __shfl_xor_sync(FULL, key_hi, d, G);
                key_hi = o > key_hi ? o : key_hi;
            }
            if (lig == 0 && valid && sweep_valid) {
                if (id_lo != 0xffffffffu) {
                    AlignEnd *e = ep.out + (reverse ? (size_t)g_lo : (size_t)id_lo * p.n_cseq + cj);
                    int b = (int)(key_lo >> 40);
                    e->best = (PACKED && b >= p.ovf_thresh) ? -1 : b;
                    e->r_end = 0xFFFFFu - (uint32_t)((key_lo >> 20) & 0xFFFFFu);
                    e->c_end = 0xFFFFFu - (uint32_t)(key_lo & 0xFFFFFu);
                }
                if (PACKED && id_hi != 0xffffffffu) {
                    AlignEnd *e = ep.out + (reverse ? (size_t)g_hi : (size_t)id_hi * p.n_cseq + cj);
                    int b = (int)(key_hi >> 40);
                    e->best = (b >= p.ovf_thresh) ? -1 : b;
                    e->r_end = 0xFFFFFu - (uint32_t)((key_hi >> 20) & 0xFFFFFu);
                    e->c_end = 0xFFFFFu - (uint32_t)(key_hi & 0xFFFFFu);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// between the passes: status / tier from the forward score, bucket the mapped pairs by (profiled, c_end)
// ---------------------------------------------------------------------------------------------
struct RangesParams {
    AlignEnd *ends;           // forward results; aux = bucket key or 0xffffffff
    const AlignEnd *starts;   // reverse results (finalize)
    const uint64_t *roff;
    uint32_t n_cseq, chunk_first, n_slots;
    uint32_t key_stride;      // max profiled length (key = cj * key_stride + c_end)
    uint32_t *hist;
    uint32_t *score;
    uint8_t *status, *tier;
    uint32_t *ref_start, *ref_end, *query_start, *query_end;
    unsigned long long *counters;  // [7] forward/reverse score mismatches (internal check)
    int invert;               // 1: SeqSrc::Query(streamed): the ranges swap sides (alignment/mod.rs:176-190)
    TierPolicy tp;
};

__global__ void ranges_classify_kernel(const RangesParams t) {
    const uint32_t pairs = t.n_slots * t.n_cseq;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pairs) return;
    const uint32_t seq = t.chunk_first + k / t.n_cseq, cj = k % t.n_cseq;
    const size_t gid = (size_t)seq * t.n_cseq + cj;
    const AlignEnd e = t.ends[gid];
    const uint32_t n = (uint32_t)(t.roff[seq + 1] - t.roff[seq]);
    t.ref_start[gid] = t.ref_end[gid] = t.query_start[gid] = t.query_end[gid] = 0;
    t.ends[gid].aux = 0xffffffffu;
    const uint8_t tier = e.best > 0 ? tier_for(t.tp, (uint32_t)e.best) : t.tp.first;
    if (e.best <= 0 || n == 0) {
        t.score[gid] = 0;
        t.status[gid] = 2;  // Unmapped (striped.rs:225-227, 608-633)
        t.tier[gid] = t.tp.first;
        return;
    }
    if (tier == 0) {
        t.score[gid] = 0;
        t.status[gid] = 1;  // Overflowed in every allowed integer type
        t.tier[gid] = t.tp.last;
        return;
    }
    t.score[gid] = (uint32_t)e.best;
    t.status[gid] = 0;
    t.tier[gid] = tier;
    const uint32_t key = cj * t.key_stride + e.c_end;
    t.ends[gid].aux = key;
    atomicAdd(&t.hist[key], 1u);
}

__global__ void ranges_scatter_kernel(const RangesParams t, const uint32_t *bucket_start, uint32_t *items) {
    const uint32_t pairs = t.n_slots * t.n_cseq;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pairs) return;
    const uint32_t seq = t.chunk_first + k / t.n_cseq, cj = k % t.n_cseq;
    const size_t gid = (size_t)seq * t.n_cseq + cj;
    const uint32_t key = t.ends[gid].aux;
    if (key == 0xffffffffu) return;
    items[bucket_start[key] + atomicAdd(&t.hist[key], 1u)] = (uint32_t)gid;
}

// ranges from the two passes: streamed rows [r_end - r', r_end + 1), profiled columns [c_end - c', c_end + 1)
__global__ void ranges_finalize_kernel(const RangesParams t) {
    const uint32_t pairs = t.n_slots * t.n_cseq;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pairs) return;
    const uint32_t seq = t.chunk_first + k / t.n_cseq, cj = k % t.n_cseq;
    const size_t gid = (size_t)seq * t.n_cseq + cj;
    const AlignEnd e = t.ends[gid];
    if (e.aux == 0xffffffffu) return;
    const AlignEnd s = t.starts[gid];
    if (s.best != e.best) atomicAdd(&t.counters[7], 1ULL);  // debug_assert_eq!(score, score2), striped.rs:381
    const uint32_t r0 = e.r_end - s.r_end, r1 = e.r_end + 1, c0 = e.c_end - s.c_end, c1 = e.c_end + 1;
    // un-inverted: the streamed sequence is zoe's `reference` argument, the profiled one its query
    t.ref_start[gid] = t.invert ? c0 : r0;
    t.ref_end[gid] = t.invert ? c1 : r1;
    t.query_start[gid] = t.invert ? r0 : c0;
    t.query_end[gid] = t.invert ? r1 : c1;
}

}  // namespace zoe_cuda
