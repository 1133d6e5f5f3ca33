// sw_ends_long.cuh -- score + exact end cell for long streamed sequences (rows > 1024), forward and reverse (sm_100a).
//
// zoe: sw_simd_score_ends (src/alignment/sw/striped.rs:153-162, 213-336) and the reverse pass of
// sw_simd_score_ranges (striped.rs:355-388) have no length limit; this is their long-row counterpart, built on the
// chunked sweep of sw_score_long.cuh (one warp per task, rows cut into chunks of 32 lanes x K rows, the H / E leaving
// a chunk parked in a boundary row in global memory).  What it adds:
//
//   * every chunk keeps its OWN boundary row (chunk c's bottom row = chunk c+1's top row stays readable), so any chunk
//     can be swept again later;
//   * branch-free bookkeeping per lane, once per two columns (as in sw_align_scan_kernel): running maximum and the first
//     / last step pair in which the lane reached it.  After a chunk, the lowest lane holding the chunk's maximum is
//     recorded; a later chunk replaces the record only with a strictly larger maximum -- zoe wants the smallest row
//     among the cells holding the best score, then the smallest column (striped.rs:305-334), and chunks / lanes own
//     ascending row ranges;
//   * a PIN sweep of the one chunk that holds the best cell: the plain recurrence again from that chunk's top boundary
//     row up to the last step in which the winning lane saw the maximum; the winning lane looks at its K rows in the
//     steps between first and last occurrence and takes (min row, then min column).
//   * REV: rows = the reversed prefix streamed[r_end], streamed[r_end-1], ..., columns = profiled[c_end], profiled[c_end-1],
//     ... bounded by the score exactly as in sw_align_scan_kernel<.., REV>; one item per task (the high half idles:
//     two long reads almost never share an end column).
//
// PACKED = two sequences in the 16-bit halves (exact while every H < 32767 - max weight - gap_open; the host checks the
// static bound and uses the 32-bit instantiation otherwise).
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "sw_align.cuh"
#include "sw_score.cuh"

namespace zoe_cuda {

struct EndsLongParams {
    ScoreParams s;            // forward: s.task_ids = length-sorted sequence ids, s.n_rseq = list length, s.n_tasks
    uint2 *boundary;          // [warp slots][chunk_cap][max_L]
    uint32_t max_L;           // columns per boundary row
    uint32_t chunk_cap;       // boundary rows per warp slot (>= chunks of the longest sequence)
    unsigned int *queue;      // dynamic task counter (zeroed before launch)
    const uint32_t *n_tasks_dev;  // optional device-side task count (REV: the item list is built on the device)
    AlignEnd *out;            // forward: [seq * n_cseq + cj]; reverse: [global pair id]
    const uint32_t *items;    // REV: global pair ids, one per task
    const AlignEnd *rev_in;   // REV: forward end cells
    int rev_maxw;             // REV: largest substitution weight (bounds the columns a reversed sub-problem needs)
    unsigned long long *counters;  // [7] scan and pin disagree (internal error)
};

#ifndef ZOE_ENDS_LONG_THREADS
#define ZOE_ENDS_LONG_THREADS 384
#endif
template <int K, bool PACKED, bool REV>
__global__ void __launch_bounds__(ZOE_ENDS_LONG_THREADS) sw_ends_long_kernel(const EndsLongParams lp) {
    using O = Ops<PACKED>;
    const ScoreParams &p = lp.s;
    constexpr int G = 32;
    constexpr int K4 = K / 4;
    constexpr int NH = PACKED ? 2 : 1;  // sequences per task
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
    if (p.cols_in_smem) stage_with_tma(s_cc, p.ccodes, p.ccodes_bytes);
    __syncthreads();
    const uint8_t *cc = p.cols_in_smem ? s_cc : p.ccodes;

    uint32_t go_s = O::splat(p.go), neg_ge = O::splat(-p.ge);
    const uint32_t one_s = O::splat(1);
    uint32_t tab_lane_off = (uint32_t)warp * (uint32_t)tab_bytes + (uint32_t)lane * 16u;
    uint32_t nz = lane != 0 ? 1u : 0u;
    asm volatile("" : "+r"(go_s), "+r"(neg_ge), "+r"(tab_lane_off), "+r"(nz));
    uint2 *bnd_slot = lp.boundary + (size_t)(blockIdx.x * warps_per_block + warp) * lp.chunk_cap * lp.max_L;

    // value of half h of a packed register (PACKED) / the register (32-bit mode)
    auto half_s = [](uint32_t v, int h) -> int { return PACKED ? (int)(int16_t)((v >> (16 * h)) & 0xffffu) : (int)v; };
    auto half_u = [](uint32_t v, int h) -> uint32_t { return PACKED ? ((v >> (16 * h)) & 0xffffu) : v; };

    const uint32_t n_tasks = lp.n_tasks_dev ? *lp.n_tasks_dev : p.n_tasks;
    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(lp.queue, 1u);
        task = __shfl_sync(FULL, task, 0);
        if (task >= n_tasks) break;

        // ---- the task's sequences ----
        uint32_t id[2] = {0xffffffffu, 0xffffffffu};   // forward: sequence ids; REV: global pair id in id[0]
        int64_t base[2] = {0, 0};                       // byte offset of row 0
        int len[2] = {0, 0};
        uint32_t rev_cj = 0;
        int rev_cend = 0, rev_cols = 0;
        if (!REV) {
            if (PACKED) {
                id[0] = p.task_ids[2 * task];
                if (2 * task + 1 < p.n_rseq) id[1] = p.task_ids[2 * task + 1];
            } else {
                id[0] = p.task_ids[task];
            }
#pragma unroll
            for (int h = 0; h < NH; ++h)
                if (id[h] != 0xffffffffu) {
                    base[h] = (int64_t)p.roff[id[h]];
                    len[h] = (int)(p.roff[id[h] + 1] - p.roff[id[h]]);
                }
        } else {
            const uint32_t gid = lp.items[task];
            const AlignEnd e = lp.rev_in[gid];
            const uint32_t seq = gid / p.n_cseq;
            id[0] = gid;
            rev_cj = gid % p.n_cseq;
            rev_cend = (int)e.c_end;
            base[0] = (int64_t)p.roff[seq] + e.r_end;  // row r' reads streamed[r_end - r']
            len[0] = (int)e.r_end + 1;
            // columns a reversed sub-problem can need: see sw_align_scan_kernel (sw_align_win.cuh), rev_bound
            long long w = 0x7fffffffLL;
            if (p.ge > 0) {
                const long long nr = (long long)e.r_end + 1;
                const long long num = nr * (long long)lp.rev_maxw - (long long)p.go - (long long)e.best;
                const long long ic = num >= 0 ? 1 + num / (long long)p.ge : 0;
                w = nr + ic + 1;
            }
            rev_cols = w > 0x7fffffffLL ? 0x7fffffff : (int)w;
        }
        const int nmax = max(len[0], len[1]);
        const int R = G * K;
        const int nchunks = max(1, (nmax + R - 1) / R);

        for (uint32_t sweep_i = 0; sweep_i < (REV ? 1u : p.n_cseq); ++sweep_i) {
            const uint32_t cj = REV ? rev_cj : sweep_i;
            const uint32_t c0 = p.coff[cj];
            const int L = REV ? min(rev_cend + 1, rev_cols) : (int)(p.coff[cj + 1] - c0);
            const uint8_t *cs = cc + c0 + (REV ? rev_cend : 0);    // column j reads cs[j] (REV: cs[-j])
            const uint8_t *scs = s_cc + c0 + (REV ? rev_cend : 0);  // the same through a pointer known to be shared

            // best cell record per sequence of the task
            int g_best[2] = {0, 0}, g_chunk[2] = {0, 0}, g_lane[2] = {0, 0};
            uint32_t g_first[2] = {0, 0}, g_last[2] = {0, 0};  // step pairs
            int bi[2] = {K, K}, bj[2] = {0, 0};                 // pin result: row inside the winning lane, column

            // One sweep of chunk `ch` over the columns.  PIN = false: all columns, bookkeeping, bottom boundary row
            // written.  PIN = true: steps [0, pin_steps), nothing written, the winning lanes search their rows.
            auto sweep_chunk = [&](auto pin_c, const int ch, const int pin_steps) {
                constexpr bool PIN = decltype(pin_c)::value;
                const int row0 = ch * R;
                const int rows_left = nmax - row0;
                const int k4_eff = min(K4, max(1, (rows_left + 4 * G - 1) / (4 * G)));
                const int k_eff = 4 * k4_eff;
                const bool has_top = ch > 0, has_bottom = !PIN && ch + 1 < nchunks;
                uint2 *bnd_top = bnd_slot + (size_t)(ch > 0 ? ch - 1 : 0) * lp.max_L;
                uint2 *bnd_bot = bnd_slot + (size_t)ch * lp.max_L;

                __syncwarp();
                for (int s = 0; s < p.n_csym; ++s) {
                    const int8_t *wrow = s_wk + s * p.S;
                    for (int i4 = 0; i4 < k4_eff; ++i4) {
                        uint32_t w[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int r = row0 + lane * k_eff + i4 * 4 + q;
                            int wl = kPadWeight, wh = kPadWeight;
                            if (r < len[0]) wl = wrow[s_lut[p.rseq[base[0] + (REV ? -(int64_t)r : (int64_t)r)]]];
                            if (PACKED && r < len[1]) wh = wrow[s_lut[p.rseq[base[1] + r]]];
                            w[q] = PACKED ? ((uint32_t)(wl & 0xffff) | ((uint32_t)(wh & 0xffff) << 16)) : (uint32_t)wl;
                        }
                        tab[(s * K4 + i4) * G + lane] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                __syncwarp();

                uint32_t H[2][K], F[K];
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    H[0][i] = H[1][i] = 0;
                    F[i] = 0;
                }
                uint32_t h_last = 0, e_out = 0, h_up_prev = 0;
                uint32_t bestp = 0, firstp = 0, lastp = 0, cm = 0;
                uint2 cur = make_uint2(0, 0), nxt = make_uint2(0, 0);
                if (has_top && lane < L) cur = __ldcg(bnd_top + lane);
                const int nsteps = PIN ? min(pin_steps, L + G - 1) : L + G - 1;
                const uint32_t top_on = has_top ? 1u - nz : 0u;

                auto column = [&](auto parity, auto fullk, const int j, const uint32_t h_in, const uint32_t e_in,
                                  const uint32_t code) {
                    constexpr int PO = decltype(parity)::value, PN = 1 - PO;
                    constexpr bool FULLK = decltype(fullk)::value;
                    const uint4 *tp = reinterpret_cast<const uint4 *>(smem + tab_lane_off + code * (uint32_t)(K4 * G * 16));
                    uint32_t diag = h_up_prev;
                    uint32_t E = e_in;
#pragma unroll
                    for (int i4 = 0; i4 < K4; ++i4) {
                        if (FULLK || i4 < k4_eff) {
                            const uint4 w4 = tp[i4 * G];
                            const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
                            uint32_t hp = 0;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int i = i4 * 4 + q;
                                const uint32_t x = O::max3(E, F[i], go_s) - go_s;
                                const uint32_t Hn = O::addmax(diag, w[q], x);
                                diag = H[PO][i];
                                E = O::addmax(E, neg_ge, Hn);
                                F[i] = O::addmax(F[i], neg_ge, Hn);
                                H[PN][i] = Hn;
                                if (!PIN) {
                                    if (q & 1)
                                        cm = O::max3(cm, Hn, hp);
                                    else
                                        hp = Hn;
                                }
                            }
                            h_last = H[PN][i4 * 4 + 3];
                        }
                    }
                    e_out = E;
                    if (has_bottom && lane == G - 1) bnd_bot[j] = make_uint2(h_last, e_out);
                };
                // m - bestp and m - cm are >= 0 in every half, so the plain 32-bit subtractions are exact
                auto bookkeeping = [&](const uint32_t s_even) {
                    const uint32_t pp = PACKED ? (s_even >> 1) * 0x00010001u : (s_even >> 1);
                    const uint32_t m = O::max2(cm, bestp);
                    const uint32_t inc = O::min2(m - bestp, one_s) * (PACKED ? 0xffffu : 0xffffffffu);
                    const uint32_t ge = (one_s - O::min2(m - cm, one_s)) * (PACKED ? 0xffffu : 0xffffffffu);
                    firstp = (firstp & ~inc) | (pp & inc);
                    lastp = (lastp & ~ge) | (pp & ge);
                    bestp = m;
                    cm = 0;
                };
                auto inputs = [&](const int step, uint32_t &h_in, uint32_t &e_in) {
                    if (has_top && (step & 31) == 0) {
                        const int idx = step + 32 + lane;
                        nxt = (idx < L) ? __ldcg(bnd_top + idx) : make_uint2(0, 0);
                    }
                    const uint32_t h_sh = __shfl_up_sync(FULL, h_last, 1), e_sh = __shfl_up_sync(FULL, e_out, 1);
                    const uint32_t h_top = __shfl_sync(FULL, cur.x, step & 31), e_top = __shfl_sync(FULL, cur.y, step & 31);
                    h_in = h_sh * nz + h_top * top_on;
                    e_in = e_sh * nz + e_top * top_on;
                    if ((step & 31) == 31) cur = nxt;
                };
                using P0 = std::integral_constant<int, 0>;
                using P1 = std::integral_constant<int, 1>;
                // PIN: the winning lane of sequence h looks for the best value among its rows (smallest row, then
                // smallest column: steps ascend, a later step only wins with a strictly smaller row)
                // (behind a warp-uniform branch: as predicated straight-line code the search ran in every step)
                auto pin_search = [&](auto parity, const int step, const int j, const bool in_cols) {
                    constexpr int PN = 1 - decltype(parity)::value;
                    bool hit[NH], any = false;
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        hit[h] = in_cols && g_best[h] > 0 && g_chunk[h] == ch && lane == g_lane[h] &&
                                 (uint32_t)step >= 2u * g_first[h] && (uint32_t)step <= 2u * g_last[h] + 1u;
                        any = any || hit[h];
                    }
                    if (!__any_sync(FULL, any)) return;
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        if (hit[h]) {
                            int irow = K;
#pragma unroll
                            for (int i = K - 1; i >= 0; --i)
                                if (i < k_eff && half_s(H[PN][i], h) == g_best[h]) irow = i;
                            if (irow < bi[h]) {
                                bi[h] = irow;
                                bj[h] = j;
                            }
                        }
                    }
                };
                auto generic_step = [&](auto parity, const int step) {
                    uint32_t h_in, e_in;
                    inputs(step, h_in, e_in);
                    const int j = step - lane;
                    const bool in_cols = j >= 0 && j < L;
                    if (in_cols) {
                        column(parity, std::false_type{}, j, h_in, e_in, (uint32_t)cs[REV ? -j : j]);
                        if (!PIN) bookkeeping((uint32_t)step & ~1u);
                    }
                    if (PIN) pin_search(parity, step, j, in_cols);
                    h_up_prev = h_in;
                };
                auto steady_step = [&](auto parity, auto fullk, const int step) {
                    uint32_t h_in, e_in;
                    inputs(step, h_in, e_in);
                    const int j = step - lane;
                    column(parity, fullk, j, h_in, e_in, (uint32_t)scs[REV ? -j : j]);
                    h_up_prev = h_in;
                };
                auto block_step = [&](auto parity, const int step, const int k) {  // inside a 32-step block: no per-step tests
                    const uint32_t h_sh = __shfl_up_sync(FULL, h_last, 1), e_sh = __shfl_up_sync(FULL, e_out, 1);
                    const uint32_t h_top = __shfl_sync(FULL, cur.x, k), e_top = __shfl_sync(FULL, cur.y, k);
                    const uint32_t h_in = h_sh * nz + h_top * top_on, e_in = e_sh * nz + e_top * top_on;
                    const int j = step - lane;
                    column(parity, std::true_type{}, j, h_in, e_in, (uint32_t)scs[REV ? -j : j]);
                    h_up_prev = h_in;
                };

                const bool fast = !PIN && p.cols_in_smem != 0;
                const bool fullk = k4_eff == K4;
                int seg_end[3] = {fast ? min(G, nsteps) : nsteps, fast ? L - 1 : 0, nsteps};
                int step = 0;
#pragma unroll 1
                for (int seg = 0; seg < 3; ++seg) {
                    const int end = seg_end[seg];
                    if (seg == 1) {
                        if (fast && fullk) {
                            // whole blocks of 32 steps: boundary-row prefetch and hand-over once per block (sw_score_long.cuh)
                            for (; step + 32 <= end; step += 32) {
                                if (has_top) {
                                    const int idx = step + 32 + lane;
                                    nxt = (idx < L) ? __ldcg(bnd_top + idx) : make_uint2(0, 0);
                                }
#pragma unroll 1
                                for (int k = 0; k < 32; k += 2) {
                                    block_step(P0{}, step + k, k);
                                    block_step(P1{}, step + k + 1, k + 1);
                                    bookkeeping((uint32_t)(step + k));
                                }
                                cur = nxt;
                            }
                            for (; step < end; step += 2) {
                                steady_step(P0{}, std::true_type{}, step);
                                steady_step(P1{}, std::true_type{}, step + 1);
                                bookkeeping((uint32_t)step);
                            }
                        } else if (fast) {
                            for (; step < end; step += 2) {
                                steady_step(P0{}, std::false_type{}, step);
                                steady_step(P1{}, std::false_type{}, step + 1);
                                bookkeeping((uint32_t)step);
                            }
                        }
                    } else {
                        for (; step < end; ++step) {
                            if (step & 1)
                                generic_step(P1{}, step);
                            else
                                generic_step(P0{}, step);
                        }
                    }
                }
                __syncwarp();
                if (!PIN) {
                    // ---- the chunk's maximum per sequence; the lowest lane holding it (smallest rows) ----
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        const int mine = half_s(bestp, h);
                        int bmax = mine;
#pragma unroll
                        for (int d = 16; d >= 1; d >>= 1) bmax = max(bmax, __shfl_xor_sync(FULL, bmax, d));
                        const unsigned holders = __ballot_sync(FULL, mine == bmax);
                        const int wl = __ffs(holders) - 1;
                        const uint32_t f = half_u(__shfl_sync(FULL, firstp, wl), h), l = half_u(__shfl_sync(FULL, lastp, wl), h);
                        if (bmax > g_best[h]) {
                            g_best[h] = bmax;
                            g_chunk[h] = ch;
                            g_lane[h] = wl;
                            g_first[h] = f;
                            g_last[h] = l;
                        }
                    }
                }
                return k_eff;
            };

            for (int ch = 0; ch < nchunks; ++ch) sweep_chunk(std::false_type{}, ch, 0);

            // ---- pin: one more sweep of the chunk(s) holding the best cells ----
            int keff_of[2] = {K, K};
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                if (g_best[h] <= 0) continue;
                if (h == 1 && g_best[0] > 0 && g_chunk[1] == g_chunk[0]) {
                    keff_of[1] = keff_of[0];  // searched together with sequence 0
                    continue;
                }
                int last_step = (int)(2u * g_last[h] + 1u);
                if (PACKED && h == 0 && g_best[1] > 0 && g_chunk[1] == g_chunk[0]) last_step = max(last_step, (int)(2u * g_last[1] + 1u));
                keff_of[h] = sweep_chunk(std::true_type{}, g_chunk[h], last_step + 1);
            }

            {
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    if (id[h] == 0xffffffffu) continue;  // warp-uniform
                    // the winning lane holds (bi, bj); everyone learns them
                    const int src = g_best[h] > 0 ? g_lane[h] : 0;
                    const int rbi = __shfl_sync(FULL, bi[h], src), rbj = __shfl_sync(FULL, bj[h], src);
                    if (lane != 0) continue;
                    AlignEnd e;
                    e.best = g_best[h];
                    if (PACKED && g_best[h] >= p.ovf_thresh) e.best = -1;
                    e.r_end = 0;
                    e.c_end = 0;
                    e.aux = 0;
                    if (g_best[h] > 0) {
                        if (rbi >= K) atomicAdd(&lp.counters[7], 1ULL);  // scan and pin disagree: internal error
                        e.r_end = (uint32_t)(g_chunk[h] * R + g_lane[h] * keff_of[h] + rbi);
                        e.c_end = (uint32_t)rbj;
                    }
                    if (REV)
                        lp.out[id[0]] = e;
                    else
                        lp.out[(size_t)id[h] * p.n_cseq + cj] = e;
                }
            }
        }
    }
}

// Between the passes (long-row path): status / tier / score from the forward result, and the list of mapped pairs for the
// reverse pass -- visited in the forward pass's longest-first order, so the reverse queue hands out the big items first.
struct RangesLongParams {
    AlignEnd *ends;
    const uint64_t *roff;
    const uint32_t *order;    // sorted sequence ids
    uint32_t n_seq, n_cseq;
    uint32_t *score;
    uint8_t *status, *tier;
    uint32_t *ref_start, *ref_end, *query_start, *query_end;
    uint32_t *items;
    uint32_t *n_items;
    TierPolicy tp;
};

__global__ void ranges_long_classify_kernel(const RangesLongParams t) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= t.n_seq * t.n_cseq) return;
    const uint32_t seq = t.order[k / t.n_cseq], cj = k % t.n_cseq;
    const size_t gid = (size_t)seq * t.n_cseq + cj;
    const AlignEnd e = t.ends[gid];
    const uint32_t n = (uint32_t)(t.roff[seq + 1] - t.roff[seq]);
    t.ref_start[gid] = t.ref_end[gid] = t.query_start[gid] = t.query_end[gid] = 0;
    t.ends[gid].aux = 0xffffffffu;
    const uint8_t tier = e.best > 0 ? tier_for(t.tp, (uint32_t)e.best) : t.tp.first;
    if (e.best <= 0 || n == 0) {
        t.score[gid] = 0;
        t.status[gid] = 2;  // Unmapped
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
    t.ends[gid].aux = 0;
    t.items[atomicAdd(t.n_items, 1u)] = (uint32_t)gid;
}

}  // namespace zoe_cuda
