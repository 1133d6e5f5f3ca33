// sw_3pass.cuh -- third pass of zoe's 3-pass local alignment (src/alignment/sw/three_pass.rs:21-104): given the
// score and the bounding box from the ranges pipeline (sw_ranges.cuh = sw_simd_score_ranges, striped.rs:355-388),
// produce the CIGAR without a full traceback matrix:
//   * no-gaps shortcut (three_pass.rs:39-59): equal range lengths and the diagonal's weights add up to the score;
//   * otherwise sw_banded_align (src/alignment/sw/banded.rs:40-133) on the box with band width |dr - dq| + 1,
//     doubled until the banded score equals the pass-1 score (three_pass.rs:66-79);
//   * otherwise sw_scalar_align on the box (three_pass.rs:81-84, src/alignment/sw/scalar.rs:173-271).
// One thread owns one pair and runs zoe's loops literally (the recurrences are tiny: a box of 150 x ~150 cells with a
// band of 5..9 columns), so tie-breaks, flag bytes and the walk are zoe's by construction.  The work is < 3 % of the two
// ranges passes that precede it.
#pragma once
#include "sw_align.cuh"

namespace zoe_cuda {

constexpr uint32_t kTpNoDp = 0xffffffffu;    // pair emits no CIGAR (not Some)
constexpr uint32_t kTpNoGaps = 0xfffffffeu;  // pair took the no-gaps shortcut: its CIGAR follows from the ranges

struct ThreePassParams {
    const uint8_t *rseq;        // streamed batch, raw bytes
    const uint64_t *roff;
    const uint8_t *pbytes;      // profiled sequences, raw bytes
    const uint32_t *coff;
    uint32_t n_cseq;
    const int8_t *weights;      // S*S, zoe's weights[ref_idx][query_idx]
    int S;
    const uint8_t *lut;
    int go, ge;                 // zoe's signed gap weights (<= 0)
    int invert;                 // 1: SeqSrc::Query(streamed) -> Alignment::invert (output.rs:396-425)
    uint64_t pair_first;        // chunk = global pair ids [pair_first, pair_first + n_pairs)
    uint32_t n_pairs;
    uint32_t *score;            // per pair, from the ranges pipeline (final orientation)
    uint8_t *status;
    uint32_t *ref_start, *ref_end, *query_start, *query_end;
    uint32_t *cig_count;
    uint32_t *dp_pair;          // [n_pairs] slot -> global pair id
    unsigned long long *dp_off; // [n_pairs] slot -> byte offset of the pair's scratch inside `blob`
    uint32_t *dp_cap;           // [n_pairs] slot -> CIGAR scratch capacity (words) at the head of the pair's scratch
    uint32_t *dp_slot;          // [n_pairs] chunk-local pair -> slot, kTpNoDp or kTpNoGaps
    uint8_t *blob;
    unsigned long long *ctr;    // per classification: [0] DP pairs, [1] scratch bytes, [5] no-gaps pairs; per call:
                                // [6] banded, [7] scalar fallbacks, [8] banded attempts, [9] running CIGAR base,
                                // [10] walks outside the band storage (zoe would panic), [11] CIGAR scratch overflows
};

__host__ __device__ inline unsigned long long tp_align16(unsigned long long x) { return (x + 15ull) & ~15ull; }
__host__ __device__ inline uint32_t tp_cig_cap(uint32_t rn, uint32_t qn) { return 2u * (rn < qn ? rn : qn) + 8u; }

struct TpBox {
    uint32_t seq, cj;
    uint32_t n, m;            // streamed / profiled lengths
    uint32_t rs, re, qs, qe;  // un-inverted: rows (streamed) [rs,re), columns (profiled) [qs,qe)
};

__device__ inline TpBox tp_box(const ThreePassParams &t, uint64_t gid) {
    TpBox b;
    b.seq = (uint32_t)(gid / t.n_cseq);
    b.cj = (uint32_t)(gid % t.n_cseq);
    b.n = (uint32_t)(t.roff[b.seq + 1] - t.roff[b.seq]);
    b.m = t.coff[b.cj + 1] - t.coff[b.cj];
    if (t.invert) {  // ranges_finalize_kernel swapped them
        b.rs = t.query_start[gid];
        b.re = t.query_end[gid];
        b.qs = t.ref_start[gid];
        b.qe = t.ref_end[gid];
    } else {
        b.rs = t.ref_start[gid];
        b.re = t.ref_end[gid];
        b.qs = t.query_start[gid];
        b.qe = t.query_end[gid];
    }
    return b;
}

// Pass 3a: the no-gaps shortcut, and scratch allocation for the pairs that need a DP.
__global__ void __launch_bounds__(256) tp_classify_kernel(const ThreePassParams t) {
    __shared__ uint8_t s_lut[256];
    __shared__ int8_t s_w[32 * 32];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = t.lut[i];
    for (int i = threadIdx.x; i < t.S * t.S; i += blockDim.x) s_w[i] = t.weights[i];
    __syncthreads();
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= t.n_pairs) return;
    const uint64_t gid = t.pair_first + k;
    t.cig_count[gid] = 0;
    t.dp_slot[k] = kTpNoDp;
    if (t.status[gid] != 0) return;
    const TpBox b = tp_box(t, gid);
    if (b.qe <= b.qs) {  // query_range.is_empty() -> Unmapped (three_pass.rs:37-38)
        t.status[gid] = 2;
        t.score[gid] = 0;
        t.ref_start[gid] = t.ref_end[gid] = t.query_start[gid] = t.query_end[gid] = 0;
        return;
    }
    const uint32_t qn = b.qe - b.qs, rn = b.re - b.rs;
    if (qn == rn) {
        const uint8_t *R = t.rseq + t.roff[b.seq] + b.rs;
        const uint8_t *P = t.pbytes + t.coff[b.cj] + b.qs;
        long long sum = 0;
        for (uint32_t i = 0; i < qn; ++i) sum += s_w[(int)s_lut[R[i]] * t.S + s_lut[P[i]]];
        if (sum >= 0 && (unsigned long long)sum == (unsigned long long)t.score[gid]) {
            // AlignmentStates::new_no_gaps (state.rs:201-208); inverted: S(ref_start) M S(ref_len - ref_end)
            const uint32_t lead = t.invert ? b.rs : b.qs, trail = t.invert ? (b.n - b.re) : (b.m - b.qe);
            t.cig_count[gid] = 1u + (lead ? 1u : 0u) + (trail ? 1u : 0u);
            t.dp_slot[k] = kTpNoGaps;
            atomicAdd(&t.ctr[5], 1ULL);
            return;
        }
    }
    // scratch of a DP pair: [CIGAR, back-filled: cap words][h_row/e_row: qn x int2][flag bytes: rn x qn]
    // (a band of width 2 bw + 1 <= qn always fits the full box)
    const uint32_t cap = tp_cig_cap(rn, qn);
    const unsigned long long need = tp_align16(4ull * cap) + tp_align16(8ull * qn) + tp_align16((unsigned long long)rn * qn);
    const uint32_t slot = (uint32_t)atomicAdd(&t.ctr[0], 1ULL);
    t.dp_pair[slot] = (uint32_t)(gid - t.pair_first);
    t.dp_cap[slot] = cap;
    t.dp_off[slot] = atomicAdd(&t.ctr[1], need);
    t.dp_slot[k] = slot;
}

// zoe's scalar / banded recurrence over the box, flags into `fl`.  BANDED: src/alignment/sw/banded.rs:55-125 (rows keep
// the columns [r - bw, r + bw], flag row stride 2 bw + 1); otherwise src/alignment/sw/scalar.rs:190-262.
// Returns the best score; (r_end, c_end) = its first occurrence in row-major order; rows_done = rows the loop visited.
// RING (banded, 2 bw + 2 <= kTpRing): h_row / e_row live in a ring of kTpRing columns in shared memory
// (he[(c % kTpRing) * kTpThreads], `he` already offset by the thread index).  A row only touches the columns of its band,
// the band moves right by one column per row, and a column that left the band is never read again, so the ring holds
// everything zoe's full-length vectors would be asked for; the column entering the band is (re)initialised to zoe's
// initial (0, gap_open) first.  Shared-memory latency instead of an L2 round trip per cell: 4.4 -> 1 ms on config 3.
constexpr uint32_t kTpRing = 32, kTpThreads = 64;

template <bool BANDED, bool RING = false>
__device__ inline int32_t tp_fill(const uint8_t *R, uint32_t rn, const uint8_t *P, uint32_t qn, const uint8_t *s_lut,
                                  const int8_t *s_w, int S, int32_t go, int32_t ge, uint32_t bw, int2 *he, uint8_t *fl,
                                  uint32_t *r_end, uint32_t *c_end, uint32_t *rows_done) {
    static_assert(!RING || BANDED, "the ring only serves the banded recurrence");
    auto HE = [&](uint32_t c) -> int2 & { return RING ? he[(c & (kTpRing - 1)) * kTpThreads] : he[c]; };
    if (RING) {
        for (uint32_t c = 0; c < min(bw, qn); ++c) HE(c) = make_int2(0, go);  // row 0's band except its entering column
    } else {
        for (uint32_t c = 0; c < qn; ++c) HE(c) = make_int2(0, go);  // h_row = 0, e_row = gap_open
    }
    int32_t best = 0, h_store = 0;
    uint32_t br = 0, bc = 0, r = 0;
    const uint32_t bfw = 2 * bw + 1;
    for (; r < rn; ++r) {
        const int8_t *wrow = s_w + (int)s_lut[R[r]] * S;
        int32_t f = go;
        int32_t h = BANDED ? h_store : 0;
        const uint32_t start_col = BANDED ? (r > bw ? r - bw : 0u) : 0u;
        const uint32_t end_col = BANDED ? min(r + bw + 1, qn) : qn;
        if (RING && r + bw < qn) HE(r + bw) = make_int2(0, go);  // the column entering the band
        if (BANDED) {
            if (start_col >= end_col) break;
            if (start_col + bw == r) h_store = max(max(h + (int32_t)wrow[s_lut[P[start_col]]], HE(start_col).y), 0);
        }
        uint8_t *frow = fl + (size_t)r * (BANDED ? bfw : qn) - (BANDED ? start_col : 0u);
        for (uint32_t c = start_col; c < end_col; ++c) {
            const int2 prev = HE(c);  // (H[r-1][c], E[r][c])
            int32_t e = prev.y;
            h += (int32_t)wrow[s_lut[P[c]]];
            h = max(max(h, e), max(f, 0));
            if (h > best) {
                best = h;
                br = r;
                bc = c;
            }
            uint32_t flag = (e == h ? 1u : 0u) | (f == h ? 4u : 0u);
            if (h == 0) flag = 16u;
            const int32_t ho = h + go;
            e = max(e + ge, ho);
            f = max(f + ge, ho);
            if (ho != go) flag |= (e > ho ? 2u : 0u) | (f > ho ? 8u : 0u);
            frow[c] = (uint8_t)flag;
            HE(c) = make_int2(h, e);
            h = prev.x;
        }
    }
    *r_end = br;
    *c_end = bc;
    *rows_done = r;
    return best;
}

// Pass 3b: one thread per pair that needs a DP.
__global__ void __launch_bounds__(kTpThreads) tp_dp_kernel(const ThreePassParams t, uint32_t n_dp) {
    __shared__ uint8_t s_lut[256];
    __shared__ int8_t s_w[32 * 32];
    __shared__ int2 s_ring[kTpRing * kTpThreads];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = t.lut[i];
    for (int i = threadIdx.x; i < t.S * t.S; i += blockDim.x) s_w[i] = t.weights[i];
    __syncthreads();
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_dp) return;
    const uint64_t gid = t.pair_first + t.dp_pair[slot];
    const TpBox b = tp_box(t, gid);
    const uint32_t qn = b.qe - b.qs, rn = b.re - b.rs;
    const uint8_t *R = t.rseq + t.roff[b.seq] + b.rs;
    const uint8_t *P = t.pbytes + t.coff[b.cj] + b.qs;
    uint8_t *base = t.blob + t.dp_off[slot];
    const uint32_t cap = t.dp_cap[slot];
    uint32_t *cig = reinterpret_cast<uint32_t *>(base);
    int2 *he = reinterpret_cast<int2 *>(base + tp_align16(4ull * cap));
    uint8_t *fl = base + tp_align16(4ull * cap) + tp_align16(8ull * qn);
    const int32_t score = (int32_t)t.score[gid];

    uint32_t r_end = 0, c_end = 0, rows_done = 0;
    uint32_t bw = (rn > qn ? rn - qn : qn - rn) + 1;
    const uint32_t max_bw = (qn - 1) / 2;
    bool banded = false;
    while (bw <= max_bw) {  // three_pass.rs:71-79
        atomicAdd(&t.ctr[8], 1ULL);
        const int32_t s = 2 * bw + 2 <= kTpRing
                              ? tp_fill<true, true>(R, rn, P, qn, s_lut, s_w, t.S, t.go, t.ge, bw, s_ring + threadIdx.x, fl, &r_end,
                                                    &c_end, &rows_done)
                              : tp_fill<true>(R, rn, P, qn, s_lut, s_w, t.S, t.go, t.ge, bw, he, fl, &r_end, &c_end, &rows_done);
        if (s > 0 && s == score) {
            banded = true;
            break;
        }
        bw *= 2;
    }
    if (!banded) {
        tp_fill<false>(R, rn, P, qn, s_lut, s_w, t.S, t.go, t.ge, 0, he, fl, &r_end, &c_end, &rows_done);
        atomicAdd(&t.ctr[7], 1ULL);
    } else {
        atomicAdd(&t.ctr[6], 1ULL);
    }

    // BackTrackable::to_alignment (backtrack.rs:290-342) over the row-major matrix (:408-411) or the band storage
    // (BandedBacktrackMatrix::move_to, :628-633: cursor = r * (2 bw + 1) + c - r.saturating_sub(bw)).
    const long long bfw = 2ll * bw + 1, band_len = (long long)rn * bfw;
    bool oob = false;
    auto cell = [&](uint32_t r, uint32_t c) -> uint32_t {
        if (!banded) return fl[(size_t)r * qn + c];
        const long long skipped = r > bw ? (long long)(r - bw) : 0;
        const long long idx = (long long)r * bfw + ((long long)c - skipped);
        if ((long long)c < skipped || idx < 0 || idx >= band_len) {
            oob = true;
            return 16u;
        }
        // cells the fill loop never visited are zero in zoe's vec![0u8; ..]
        const uint32_t rr = (uint32_t)(idx / bfw), bc = (uint32_t)(idx % bfw);
        const uint32_t sc = rr > bw ? rr - bw : 0u, ec = min(rr + bw + 1, qn);
        if (rr >= rows_done || sc >= ec || bc >= ec - sc) return 0u;
        return fl[idx];
    };

    CigarBack cg;
    cg.init(cig, cap);
    const uint32_t OP_UP = t.invert ? 1u /*I*/ : 2u /*D*/, OP_LEFT = t.invert ? 2u : 1u;
    uint32_t r = r_end + 1, c = c_end + 1;
    const uint32_t r_end1 = r, c_end1 = c;
    // un-inverted: three_pass.rs:96 soft_clip(query.len() - adjusted_query_range.end) after to_alignment's own
    // soft clip of the sub-query's tail (backtrack.rs:305); inverted: invert() drops every S and clips by the
    // (adjusted) reference range of the un-inverted alignment = the streamed rows (output.rs:399-416)
    if (t.invert) {
        cg.push(4u, b.n - (r_end1 + b.rs));
    } else {
        cg.push(4u, b.m - (c_end1 + b.qs));
        cg.push(4u, qn - c_end1);
    }
    uint32_t cur = cell(r_end, c_end);
    int op = 0;  // 0 none, 1 D(up), 2 I(left), 3 M
    while (!(cur & 16u) && r > 0 && c > 0) {
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
        cur = cell(r > 0 ? r - 1 : 0, c > 0 ? c - 1 : 0);
    }
    if (t.invert) {
        cg.push(4u, r + b.rs);
    } else {
        cg.push(4u, c);          // to_alignment's 5' soft clip of the sub-query
        cg.push(4u, c + b.qs);   // three_pass.rs:95 prepend_soft_clip(adjusted_query_range.start)
    }
    cg.flush();
    if (cg.overflow) atomicAdd(&t.ctr[11], 1ULL);
    if (oob) atomicAdd(&t.ctr[10], 1ULL);
    t.cig_count[gid] = cg.n;
    // three_pass.rs:89-105: ranges of the sub-alignment shifted into the full sequences
    const uint32_t ar0 = r + b.rs, ar1 = r_end1 + b.rs, aq0 = c + b.qs, aq1 = c_end1 + b.qs;
    t.ref_start[gid] = t.invert ? aq0 : ar0;
    t.ref_end[gid] = t.invert ? aq1 : ar1;
    t.query_start[gid] = t.invert ? ar0 : aq0;
    t.query_end[gid] = t.invert ? ar1 : aq1;
}

// Pass 3c: CIGARs into the compacted output stream (offsets from the cigar_*scan* kernels of sw_align.cuh).
__global__ void __launch_bounds__(256) tp_gather_kernel(const ThreePassParams t, const uint64_t *out_off, uint32_t *out,
                                                        uint64_t out_cap) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= t.n_pairs) return;
    const uint32_t slot = t.dp_slot[k];
    if (slot == kTpNoDp) return;
    const uint64_t gid = t.pair_first + k;
    const uint64_t o = out_off[gid];
    const uint32_t cnt = t.cig_count[gid];
    if (o + cnt > out_cap) return;  // the caller reports ZOE_CUDA_E_CIGAR_CAP from the total
    const TpBox b = tp_box(t, gid);
    if (slot == kTpNoGaps) {
        const uint32_t lead = t.invert ? b.rs : b.qs, trail = t.invert ? (b.n - b.re) : (b.m - b.qe);
        uint32_t i = 0;
        if (lead) out[o + i++] = (lead << 4) | 4u;
        out[o + i++] = ((b.qe - b.qs) << 4) | 0u;
        if (trail) out[o + i++] = (trail << 4) | 4u;
        return;
    }
    const uint32_t cap = t.dp_cap[slot];
    const uint32_t *src = reinterpret_cast<const uint32_t *>(t.blob + t.dp_off[slot]) + (cap - min(cnt, cap));
    for (uint32_t i = 0; i < min(cnt, cap); ++i) out[o + i] = src[i];
}

}  // namespace zoe_cuda
