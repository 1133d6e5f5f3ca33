/*
 * zoe_sw_cpu.cpp -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A vectorised, multi-threaded CPU restatement ("port") of zoe's striped Smith-Waterman
 * score path, used only as the reported CPU baseline of bench.py (cpu_baseline leg and
 * `--impl reference`) and cross-checked against zoe_sw_oracle.c in tests/.  zoe itself (Rust
 * nightly, portable_simd) cannot be built in this image, so this is NOT the zoe binary.
 *
 * Follows, per pair: StripedProfile::new_unchecked (src/alignment/profile.rs:270-306),
 * sw_simd_score (src/alignment/sw/striped.rs:65-142), score_to_maybe_aligned (:608-633) and the
 * i8 -> i16 -> i32 chain of ProfileSets::sw_score_from_i8 (src/alignment/profile_set.rs:71-78),
 * with zoe's lane presets (profile_set.rs:434-483).  Profiles are built once per target and
 * shared by all worker threads, like SharedProfiles (profile_set.rs:552-560).
 *
 * SIMD: GCC vector extensions (`vector_size`), with x86 saturating-add intrinsics where the
 * element width has them; std::simd's i32 saturating ops are emulated the way LLVM lowers them.
 */
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

namespace {

struct Scoring {
    const int8_t *weights;
    int S;
    const uint8_t *map;
    int gap_open, gap_extend;
};

template <typename T, int N>
struct Vec {
    typedef T type __attribute__((vector_size(N * sizeof(T))));
};

template <typename T, int N>
using V = typename Vec<T, N>::type;

template <typename T, int N>
static inline V<T, N> splat(T x) {
    V<T, N> v = {};
    for (int i = 0; i < N; i++) v[i] = x;
    return v;
}

template <typename T, int N>
static inline V<T, N> vmax(V<T, N> a, V<T, N> b) {
    return a > b ? a : b;
}

/* lane-wise saturating add / sub (src/data/extension/simd.rs:9-87) */
template <typename T, int N>
static inline V<T, N> sat_add(V<T, N> a, V<T, N> b) {
    constexpr int B = N * sizeof(T);
#if defined(__AVX512BW__)
    if constexpr (B == 64 && sizeof(T) == 1) return (V<T, N>)_mm512_adds_epi8((__m512i)a, (__m512i)b);
    if constexpr (B == 64 && sizeof(T) == 2) return (V<T, N>)_mm512_adds_epi16((__m512i)a, (__m512i)b);
#endif
#if defined(__AVX2__)
    if constexpr (B == 32 && sizeof(T) == 1) return (V<T, N>)_mm256_adds_epi8((__m256i)a, (__m256i)b);
    if constexpr (B == 32 && sizeof(T) == 2) return (V<T, N>)_mm256_adds_epi16((__m256i)a, (__m256i)b);
#endif
#if defined(__SSE2__)
    if constexpr (B == 16 && sizeof(T) == 1) return (V<T, N>)_mm_adds_epi8((__m128i)a, (__m128i)b);
    if constexpr (B == 16 && sizeof(T) == 2) return (V<T, N>)_mm_adds_epi16((__m128i)a, (__m128i)b);
#endif
    typedef typename std::make_unsigned<T>::type U;
    typedef U UV __attribute__((vector_size(N * sizeof(T))));
    V<T, N> sum = (V<T, N>)((UV)a + (UV)b);
    V<T, N> ovf = (a ^ sum) & (b ^ sum);  // sign bit set where the signed add overflowed
    V<T, N> lim = a < 0 ? splat<T, N>(std::numeric_limits<T>::min()) : splat<T, N>(std::numeric_limits<T>::max());
    return ovf < 0 ? lim : sum;
}

template <typename T, int N>
static inline V<T, N> sat_sub(V<T, N> a, V<T, N> b) {
    constexpr int B = N * sizeof(T);
#if defined(__AVX512BW__)
    if constexpr (B == 64 && sizeof(T) == 1) return (V<T, N>)_mm512_subs_epi8((__m512i)a, (__m512i)b);
    if constexpr (B == 64 && sizeof(T) == 2) return (V<T, N>)_mm512_subs_epi16((__m512i)a, (__m512i)b);
#endif
#if defined(__AVX2__)
    if constexpr (B == 32 && sizeof(T) == 1) return (V<T, N>)_mm256_subs_epi8((__m256i)a, (__m256i)b);
    if constexpr (B == 32 && sizeof(T) == 2) return (V<T, N>)_mm256_subs_epi16((__m256i)a, (__m256i)b);
#endif
#if defined(__SSE2__)
    if constexpr (B == 16 && sizeof(T) == 1) return (V<T, N>)_mm_subs_epi8((__m128i)a, (__m128i)b);
    if constexpr (B == 16 && sizeof(T) == 2) return (V<T, N>)_mm_subs_epi16((__m128i)a, (__m128i)b);
#endif
    typedef typename std::make_unsigned<T>::type U;
    typedef U UV __attribute__((vector_size(N * sizeof(T))));
    V<T, N> diff = (V<T, N>)((UV)a - (UV)b);
    V<T, N> ovf = (a ^ b) & (a ^ diff);
    V<T, N> lim = a < 0 ? splat<T, N>(std::numeric_limits<T>::min()) : splat<T, N>(std::numeric_limits<T>::max());
    return ovf < 0 ? lim : diff;
}

/* shift_elements_right::<1>(fill): lane i <- lane i-1, lane 0 <- fill.  One register permute (vpermi2b / valignr /
 * vpalignr), never through memory. */
template <typename T, int N, size_t... I>
static inline V<T, N> shr1_impl(V<T, N> v, V<T, N> f, std::index_sequence<I...>) {
    return __builtin_shufflevector(v, f, (I == 0 ? (int)N : (int)I - 1)...);
}
template <typename T, int N>
static inline V<T, N> shr1(V<T, N> v, T fill) {
    return shr1_impl<T, N>(v, splat<T, N>(fill), std::make_index_sequence<N>{});
}

template <typename T, int N>
static inline bool any_gt(V<T, N> a, V<T, N> b) {
    constexpr int B = N * sizeof(T);
#if defined(__AVX512BW__) && defined(__AVX512VL__)
    if constexpr (B == 64 && sizeof(T) == 1) return _mm512_cmpgt_epi8_mask((__m512i)a, (__m512i)b) != 0;
    if constexpr (B == 64 && sizeof(T) == 2) return _mm512_cmpgt_epi16_mask((__m512i)a, (__m512i)b) != 0;
    if constexpr (B == 64 && sizeof(T) == 4) return _mm512_cmpgt_epi32_mask((__m512i)a, (__m512i)b) != 0;
    if constexpr (B == 32 && sizeof(T) == 1) return _mm256_cmpgt_epi8_mask((__m256i)a, (__m256i)b) != 0;
    if constexpr (B == 32 && sizeof(T) == 2) return _mm256_cmpgt_epi16_mask((__m256i)a, (__m256i)b) != 0;
    if constexpr (B == 32 && sizeof(T) == 4) return _mm256_cmpgt_epi32_mask((__m256i)a, (__m256i)b) != 0;
#elif defined(__AVX2__)
    if constexpr (B == 32 && sizeof(T) == 1) return !_mm256_testz_si256(_mm256_cmpgt_epi8((__m256i)a, (__m256i)b), _mm256_set1_epi8(-1));
    if constexpr (B == 32 && sizeof(T) == 2) return !_mm256_testz_si256(_mm256_cmpgt_epi16((__m256i)a, (__m256i)b), _mm256_set1_epi8(-1));
    if constexpr (B == 32 && sizeof(T) == 4) return !_mm256_testz_si256(_mm256_cmpgt_epi32((__m256i)a, (__m256i)b), _mm256_set1_epi8(-1));
#endif
    auto m = a > b;
    constexpr int W = B / 8;
    if constexpr (W >= 1) {
        uint64_t w[W];
        std::memcpy(w, &m, sizeof(w));
        uint64_t acc = 0;
        for (int i = 0; i < W; i++) acc |= w[i];
        return acc != 0;
    } else {
        for (int i = 0; i < N; i++)
            if (m[i]) return true;
        return false;
    }
}

/* StripedProfile<T, N, S>: src/alignment/profile.rs:198-207, 270-306 */
template <typename T, int N>
struct Profile {
    std::vector<V<T, N>, std::allocator<V<T, N>>> prof;  // [S * nv]
    int nv = 0, S = 0;
    size_t seq_len = 0;
    T gap_open = 0, gap_extend = 0;

    void build(const uint8_t *seq, size_t m, const Scoring &sc) {
        nv = (int)((m + N - 1) / N);
        S = sc.S;
        seq_len = m;
        prof.assign((size_t)S * nv, splat<T, N>(0));
        size_t total = (size_t)N * nv;
        for (int v = 0; v < nv; v++)
            for (int r = 0; r < S; r++) {
                V<T, N> vec = splat<T, N>(0);
                int i = 0;
                for (size_t q = v; q < total; q += nv, i++)
                    if (q < m) vec[i] = (T)sc.weights[r * S + sc.map[seq[q]]];
                prof[(size_t)r * nv + v] = vec;
            }
        gap_open = (T)(-sc.gap_open);
        gap_extend = (T)(-sc.gap_extend);
    }
};

/* sw_simd_score: src/alignment/sw/striped.rs:65-142.  Returns 0 Some / 1 Overflowed / 2 Unmapped. */
template <typename T, int N>
static int striped_score(const Profile<T, N> &p, const uint8_t *reference, size_t n, const uint8_t *map,
                         std::vector<V<T, N>> &scratch, uint32_t *score, uint64_t *lazy_iters = nullptr) {
    uint64_t lazy = 0;
    const int nv = p.nv;
    const T MIN = std::numeric_limits<T>::min(), MAX = std::numeric_limits<T>::max();
    const V<T, N> minimums = splat<T, N>(MIN), gap_opens = splat<T, N>(p.gap_open),
                  gap_extends = splat<T, N>(p.gap_extend);
    scratch.assign((size_t)3 * nv, minimums);
    V<T, N> *load = scratch.data(), *store = load + nv, *e_scores = store + nv;
    V<T, N> max_scores = minimums;
    for (size_t r = 0; r < n; r++) {
        const V<T, N> *scores_vec = p.prof.data() + (size_t)map[reference[r]] * nv;
        V<T, N> F = minimums;
        V<T, N> H = shr1<T, N>(store[nv - 1], MIN);
        std::swap(load, store);
        for (int j = 0; j < nv; j++) {
            V<T, N> E = e_scores[j];
            H = sat_add<T, N>(H, scores_vec[j]);
            H = vmax<T, N>(vmax<T, N>(H, E), F);
            max_scores = vmax<T, N>(max_scores, H);
            store[j] = H;
            H = sat_sub<T, N>(H, gap_opens);
            E = vmax<T, N>(sat_sub<T, N>(E, gap_extends), H);
            F = vmax<T, N>(sat_sub<T, N>(F, gap_extends), H);
            e_scores[j] = E;
            H = load[j];
        }
        int j = 0;
        H = store[0];
        F = shr1<T, N>(F, MIN);
        while (any_gt<T, N>(F, sat_sub<T, N>(H, gap_opens))) {
            ++lazy;
            H = vmax<T, N>(H, F);
            store[j] = H;
            F = sat_sub<T, N>(F, gap_extends);
            if (++j >= nv) {
                j = 0;
                F = shr1<T, N>(F, MIN);
            }
            H = store[j];
        }
    }
    if (lazy_iters) *lazy_iters += lazy;
    T best = max_scores[0];
    for (int i = 1; i < N; i++) best = std::max(best, max_scores[i]);
    if (!(best < MAX)) return 1;
    uint32_t s = (uint32_t)((int64_t)MAX + 1 + (int64_t)best);
    *score = s;
    return s == 0 ? 2 : 0;
}

/* ---- alignment with traceback -------------------------------------------------------------------------------- */
constexpr uint8_t kUp = 1, kUpExt = 2, kLeft = 4, kLeftExt = 8, kStop = 16; /* backtrack.rs:18-34 */

template <typename T, int N>
static inline V<uint8_t, N> mask8(V<T, N> m) { /* Mask<T,N>::cast::<i8>() */
    return (V<uint8_t, N>)__builtin_convertvector(m, V<int8_t, N>);
}

struct AlnOut { /* Alignment<u32>, output.rs:264-279, already inverted when the streamed side is the query */
    uint32_t score, ref_start, ref_end, query_start, query_end;
    std::vector<uint32_t> cigar; /* len << 4 | op, op: 0 M, 1 I, 2 D, 4 S -- the product's CIGAR word encoding */
};

struct States { /* AlignmentStates::add_ciglet (state.rs:142-152): run-length merged, zero-length runs dropped */
    std::vector<uint32_t> w;
    void add(uint64_t inc, uint32_t op) {
        if (!inc) return;
        if (!w.empty() && (w.back() & 15u) == op)
            w.back() += (uint32_t)inc << 4;
        else
            w.push_back(((uint32_t)inc << 4) | op);
    }
};

/* sw_simd_align: src/alignment/sw/striped.rs:449-598 + BackTrackable::to_alignment (backtrack.rs:290-342, striped
 * addressing :473-477) + Alignment::invert (output.rs:396-425) when `invert`.  0 Some / 1 Overflowed / 2 Unmapped. */
template <typename T, int N>
struct AlignScratch {
    std::vector<V<T, N>> rows;
    std::vector<V<uint8_t, N>> flags;
    States st;
};

template <typename T, int N>
static int striped_align(const Profile<T, N> &p, const uint8_t *reference, size_t n, const uint8_t *map,
                         AlignScratch<T, N> &ws, bool invert, AlnOut *out) {
    if (n == 0) return 2;
    const int nv = p.nv;
    const T MIN = std::numeric_limits<T>::min(), MAX = std::numeric_limits<T>::max();
    const V<T, N> minimums = splat<T, N>(MIN), gap_opens = splat<T, N>(p.gap_open), gap_extends = splat<T, N>(p.gap_extend);
    const V<uint8_t, N> UP = splat<uint8_t, N>(kUp), UPX = splat<uint8_t, N>(kUpExt), LEFT = splat<uint8_t, N>(kLeft),
                        LEFTX = splat<uint8_t, N>(kLeftExt), STOP = splat<uint8_t, N>(kStop);
    ws.rows.assign((size_t)4 * nv, minimums);
    if (ws.flags.size() < n * (size_t)nv) ws.flags.resize(n * (size_t)nv);
    V<T, N> *load = ws.rows.data(), *store = load + nv, *e_scores = store + nv, *max_row = e_scores + nv;
    T best = MIN;
    size_t r_end = n - 1;
    for (size_t r = 0; r < n; r++) {
        V<T, N> F = minimums;
        V<T, N> H = shr1<T, N>(store[nv - 1], MIN);
        if (r > 1 && r_end == r - 2) std::swap(max_row, load);
        std::swap(load, store);
        const V<T, N> *scores_vec = p.prof.data() + (size_t)map[reference[r]] * nv;
        V<uint8_t, N> *brow = ws.flags.data() + r * (size_t)nv;
        V<T, N> max_scores = minimums;
        for (int v = 0; v < nv; v++) {
            V<T, N> E = e_scores[v];
            H = sat_add<T, N>(H, scores_vec[v]);
            H = vmax<T, N>(vmax<T, N>(H, E), F);
            V<uint8_t, N> fl = splat<uint8_t, N>(0);
            max_scores = vmax<T, N>(max_scores, H);
            fl |= mask8<T, N>(E == H) & UP;
            fl |= mask8<T, N>(F == H) & LEFT;
            const V<uint8_t, N> stopped = mask8<T, N>(H == minimums);
            store[v] = H;
            H = sat_sub<T, N>(H, gap_opens);
            E = vmax<T, N>(sat_sub<T, N>(E, gap_extends), H);
            F = vmax<T, N>(sat_sub<T, N>(F, gap_extends), H);
            fl |= mask8<T, N>(E > H) & UPX;
            fl |= mask8<T, N>(F > H) & LEFTX;
            fl = (fl & ~stopped) | (STOP & stopped);
            brow[v] = fl;
            e_scores[v] = E;
            H = load[v];
        }
        for (int k = 0; k < N; k++) { /* 'lazy_f */
            F = shr1<T, N>(F, MIN);
            bool done = false;
            for (int v = 0; v < nv; v++) {
                H = store[v];
                if (!any_gt<T, N>(F, sat_sub<T, N>(H, gap_opens))) {
                    done = true;
                    break;
                }
                H = vmax<T, N>(H, F);
                store[v] = H;
                V<uint8_t, N> fl = brow[v];
                const V<uint8_t, N> stopped = mask8<T, N>(H == minimums);
                const V<uint8_t, N> feq = mask8<T, N>(F == H);
                fl = (fl & ~feq) | (((fl & UPX) | LEFT) & feq); /* simd_correct_and_set_left */
                H = sat_sub<T, N>(H, gap_opens);
                F = sat_sub<T, N>(F, gap_extends);
                fl |= mask8<T, N>(F > H) & LEFTX;
                fl = (fl & ~stopped) | (STOP & stopped);
                brow[v] = fl;
            }
            if (done) break;
        }
        T row_best = max_scores[0];
        for (int i = 1; i < N; i++) row_best = std::max(row_best, max_scores[i]);
        if (row_best > best) {
            if (row_best >= MAX) return 1;
            best = row_best;
            r_end = r;
        }
    }
    if (r_end == n - 1)
        max_row = store;
    else if (r_end == n - 2)
        max_row = load;
    const size_t m = p.seq_len;
    size_t c_end = m - 1;
    for (size_t ci = 0; ci < m; ci++)
        if (max_row[ci % nv][ci / nv] == best) {
            c_end = ci;
            break;
        }
    if (!(best < MAX)) return 1;
    const uint32_t score = (uint32_t)((int64_t)MAX + 1 + (int64_t)best);
    if (score == 0) return 2;
    /* to_alignment */
    const uint8_t *fb = reinterpret_cast<const uint8_t *>(ws.flags.data());
    auto cell = [&](size_t r, size_t c) -> uint8_t { return fb[((size_t)nv * r + c % nv) * N + c / nv]; };
    States &st = ws.st;
    st.w.clear();
    uint32_t op = 99;
    uint8_t cur = cell(r_end, c_end);
    size_t r = r_end + 1, c = c_end + 1;
    const size_t r_end1 = r, c_end1 = c;
    st.add(m - c, 4);
    while (!(cur & kStop) && r > 0 && c > 0) {
        if (op == 2 && (cur & kUpExt)) {
            r -= 1;
        } else if (op == 1 && (cur & kLeftExt)) {
            c -= 1;
        } else if (cur & kUp) {
            op = 2;
            r -= 1;
        } else if (cur & kLeft) {
            op = 1;
            c -= 1;
        } else {
            op = 0;
            r -= 1;
            c -= 1;
        }
        st.add(1, op);
        cur = cell(r > 0 ? r - 1 : 0, c > 0 ? c - 1 : 0);
    }
    st.add(c, 4);
    std::reverse(st.w.begin(), st.w.end());
    out->score = score;
    out->cigar.clear();
    if (!invert) {
        out->ref_start = (uint32_t)r;
        out->ref_end = (uint32_t)r_end1;
        out->query_start = (uint32_t)c;
        out->query_end = (uint32_t)c_end1;
        out->cigar = st.w;
    } else { /* Alignment::invert: I <-> D, soft clips rebuilt from the (old) reference range; ciglets appended as-is */
        if (r) out->cigar.push_back(((uint32_t)r << 4) | 4u);
        for (uint32_t w : st.w) {
            const uint32_t o = w & 15u;
            if (o == 4) continue;
            out->cigar.push_back((w & ~15u) | (o == 1 ? 2u : (o == 2 ? 1u : o)));
        }
        if (n - r_end1) out->cigar.push_back(((uint32_t)(n - r_end1) << 4) | 4u);
        out->ref_start = (uint32_t)c;
        out->ref_end = (uint32_t)c_end1;
        out->query_start = (uint32_t)r;
        out->query_end = (uint32_t)r_end1;
    }
    return 0;
}

std::atomic<uint64_t> g_lazy_iters{0}, g_rows{0}; /* lazy-F vector revisits / DP rows of the last score batch */

template <int M, int N, int O>
struct ProfileSet {  // ProfileSets<M, N, O, S>: profile_set.rs:19-359
    Profile<int8_t, M> p8;
    Profile<int16_t, N> p16;
    Profile<int32_t, O> p32;
    std::once_flag f16, f32;
    const uint8_t *seq = nullptr;
    size_t m = 0;
};

template <int M, int N, int O>
static void run_batch(const uint8_t *prof_concat, const uint64_t *prof_off, uint32_t n_prof, const uint8_t *reads,
                      const uint64_t *off, uint64_t n, const Scoring &sc, int n_threads, uint32_t *score,
                      uint8_t *status, uint8_t *tier) {
    g_lazy_iters = 0;
    g_rows = 0;
    std::vector<std::unique_ptr<ProfileSet<M, N, O>>> sets;
    for (uint32_t j = 0; j < n_prof; j++) {
        auto ps = std::make_unique<ProfileSet<M, N, O>>();
        ps->seq = prof_concat + prof_off[j];
        ps->m = (size_t)(prof_off[j + 1] - prof_off[j]);
        ps->p8.build(ps->seq, ps->m, sc);
        sets.push_back(std::move(ps));
    }
    auto worker = [&](uint64_t a, uint64_t b) {
        std::vector<V<int8_t, M>> s8;
        std::vector<V<int16_t, N>> s16;
        std::vector<V<int32_t, O>> s32;
        uint64_t lazy = 0, rows = 0;
        for (uint64_t i = a; i < b; i++) {
            const uint8_t *r = reads + off[i];
            size_t len = (size_t)(off[i + 1] - off[i]);
            for (uint32_t j = 0; j < n_prof; j++) {
                ProfileSet<M, N, O> &ps = *sets[j];
                uint32_t sco = 0;
                int t = 8;
                int rc = striped_score<int8_t, M>(ps.p8, r, len, sc.map, s8, &sco, &lazy);
                rows += len;
                if (rc == 1) {  // or_else_overflowed: output.rs:81-83
                    std::call_once(ps.f16, [&] { ps.p16.build(ps.seq, ps.m, sc); });
                    t = 16;
                    rc = striped_score<int16_t, N>(ps.p16, r, len, sc.map, s16, &sco, &lazy);
                    rows += len;
                    if (rc == 1) {
                        std::call_once(ps.f32, [&] { ps.p32.build(ps.seq, ps.m, sc); });
                        t = 32;
                        rc = striped_score<int32_t, O>(ps.p32, r, len, sc.map, s32, &sco, &lazy);
                        rows += len;
                    }
                }
                size_t k = (size_t)i * n_prof + j;
                score[k] = rc == 0 ? sco : 0;
                status[k] = (uint8_t)rc;
                tier[k] = (uint8_t)t;
            }
        }
        g_lazy_iters += lazy;
        g_rows += rows;
    };
    if (n_threads <= 1) {
        worker(0, n);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(worker, n * t / n_threads, n * (t + 1) / n_threads);
    for (auto &x : th) x.join();
}

/* Batched SharedProfiles::sw_align_from_i8(SeqSrc::Query(read)) (profile_set.rs:136-179, 552-560): every read against
 * every profiled sequence, i8 -> i16 -> i32.  Per-thread CIGAR streams are concatenated in pair order afterwards. */
struct AlignBatchOut {
    uint32_t *score, *ref_start, *ref_end, *query_start, *query_end;
    uint8_t *status, *tier;
    uint32_t *cigar;
    uint64_t *cigar_off;
    uint64_t cigar_cap;
};

template <int M, int N, int O>
static int run_align_batch(const uint8_t *prof_concat, const uint64_t *prof_off, uint32_t n_prof, const uint8_t *reads,
                           const uint64_t *off, uint64_t n, const Scoring &sc, int n_threads, bool invert,
                           const AlignBatchOut &o) {
    std::vector<std::unique_ptr<ProfileSet<M, N, O>>> sets;
    for (uint32_t j = 0; j < n_prof; j++) {
        auto ps = std::make_unique<ProfileSet<M, N, O>>();
        ps->seq = prof_concat + prof_off[j];
        ps->m = (size_t)(prof_off[j + 1] - prof_off[j]);
        ps->p8.build(ps->seq, ps->m, sc);
        sets.push_back(std::move(ps));
    }
    n_threads = std::max(1, n_threads);
    std::vector<std::vector<uint32_t>> streams(n_threads);
    std::vector<uint64_t> first(n_threads + 1);
    for (int t = 0; t <= n_threads; t++) first[t] = n * t / n_threads;
    auto worker = [&](int t) {
        AlignScratch<int8_t, M> a8;
        AlignScratch<int16_t, N> a16;
        AlignScratch<int32_t, O> a32;
        AlnOut aln;
        std::vector<uint32_t> &stream = streams[t];
        for (uint64_t i = first[t]; i < first[t + 1]; i++) {
            const uint8_t *r = reads + off[i];
            const size_t len = (size_t)(off[i + 1] - off[i]);
            for (uint32_t j = 0; j < n_prof; j++) {
                ProfileSet<M, N, O> &ps = *sets[j];
                int tr = 8;
                int rc = striped_align<int8_t, M>(ps.p8, r, len, sc.map, a8, invert, &aln);
                if (rc == 1) {
                    std::call_once(ps.f16, [&] { ps.p16.build(ps.seq, ps.m, sc); });
                    tr = 16;
                    rc = striped_align<int16_t, N>(ps.p16, r, len, sc.map, a16, invert, &aln);
                    if (rc == 1) {
                        std::call_once(ps.f32, [&] { ps.p32.build(ps.seq, ps.m, sc); });
                        tr = 32;
                        rc = striped_align<int32_t, O>(ps.p32, r, len, sc.map, a32, invert, &aln);
                    }
                }
                const size_t k = (size_t)i * n_prof + j;
                o.status[k] = (uint8_t)rc;
                o.tier[k] = (uint8_t)tr;
                if (rc == 0) {
                    o.score[k] = aln.score;
                    o.ref_start[k] = aln.ref_start;
                    o.ref_end[k] = aln.ref_end;
                    o.query_start[k] = aln.query_start;
                    o.query_end[k] = aln.query_end;
                    o.cigar_off[k + 1] = aln.cigar.size(); /* length for now; prefix-summed below */
                    stream.insert(stream.end(), aln.cigar.begin(), aln.cigar.end());
                } else {
                    o.score[k] = o.ref_start[k] = o.ref_end[k] = o.query_start[k] = o.query_end[k] = 0;
                    o.cigar_off[k + 1] = 0;
                }
            }
        }
    };
    if (n_threads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(worker, t);
        for (auto &x : th) x.join();
    }
    o.cigar_off[0] = 0;
    const uint64_t pairs = n * n_prof;
    for (uint64_t k = 0; k < pairs; k++) o.cigar_off[k + 1] += o.cigar_off[k];
    if (o.cigar_off[pairs] > o.cigar_cap) return -6;
    uint64_t pos = 0;
    for (int t = 0; t < n_threads; t++) {
        if (!streams[t].empty()) std::memcpy(o.cigar + pos, streams[t].data(), streams[t].size() * sizeof(uint32_t));
        pos += streams[t].size();
    }
    return 0;
}

}  // namespace

extern "C" {

/* Batched SharedProfiles::sw_score_from_i8 on `n_threads` host threads.
 * width_bits: 128, 256 or 512 -> lane preset (profile_set.rs:434-483).  Returns 0 or -5. */
int zo_cpu_score_batch(const uint8_t *prof_concat, const uint64_t *prof_off, uint32_t n_prof, const uint8_t *reads,
                       const uint64_t *off, uint64_t n, const int8_t *weights, int S, const uint8_t *map, int gap_open,
                       int gap_extend, int width_bits, int n_threads, uint32_t *score, uint8_t *status, uint8_t *tier) {
    Scoring sc{weights, S, map, gap_open, gap_extend};
    for (uint32_t j = 0; j < n_prof; j++)
        if (prof_off[j + 1] == prof_off[j]) return -1;
    switch (width_bits) {
        case 128:
            run_batch<16, 8, 4>(prof_concat, prof_off, n_prof, reads, off, n, sc, n_threads, score, status, tier);
            return 0;
        case 256:
            run_batch<32, 16, 8>(prof_concat, prof_off, n_prof, reads, off, n, sc, n_threads, score, status, tier);
            return 0;
        case 512:
            run_batch<64, 32, 16>(prof_concat, prof_off, n_prof, reads, off, n, sc, n_threads, score, status, tier);
            return 0;
        default:
            return -5;
    }
}

/* Batched SharedProfiles::sw_align_from_i8 (SeqSrc::Query(read) when streamed_is_query) on `n_threads` host threads.
 * CIGAR words: len << 4 | op (0 M, 1 I, 2 D, 4 S); cigar_off has n * n_prof + 1 entries.  Returns 0, -5 or -6 (cap). */
int zo_cpu_align_batch(const uint8_t *prof_concat, const uint64_t *prof_off, uint32_t n_prof, const uint8_t *reads,
                       const uint64_t *off, uint64_t n, const int8_t *weights, int S, const uint8_t *map, int gap_open,
                       int gap_extend, int width_bits, int n_threads, int streamed_is_query, uint32_t *score,
                       uint8_t *status, uint8_t *tier, uint32_t *ref_start, uint32_t *ref_end, uint32_t *query_start,
                       uint32_t *query_end, uint32_t *cigar, uint64_t *cigar_off, uint64_t cigar_cap) {
    Scoring sc{weights, S, map, gap_open, gap_extend};
    for (uint32_t j = 0; j < n_prof; j++)
        if (prof_off[j + 1] == prof_off[j]) return -1;
    AlignBatchOut o{score, ref_start, ref_end, query_start, query_end, status, tier, cigar, cigar_off, cigar_cap};
    const bool inv = streamed_is_query != 0;
    switch (width_bits) {
        case 128: return run_align_batch<16, 8, 4>(prof_concat, prof_off, n_prof, reads, off, n, sc, n_threads, inv, o);
        case 256: return run_align_batch<32, 16, 8>(prof_concat, prof_off, n_prof, reads, off, n, sc, n_threads, inv, o);
        case 512: return run_align_batch<64, 32, 16>(prof_concat, prof_off, n_prof, reads, off, n, sc, n_threads, inv, o);
        default: return -5;
    }
}

/* lazy-F revisits (vectors re-processed by the correction loop) and DP rows of the last zo_cpu_score_batch call */
void zo_cpu_last_lazy_stats(uint64_t *lazy_vectors, uint64_t *rows) {
    *lazy_vectors = g_lazy_iters.load();
    *rows = g_rows.load();
}

int zo_cpu_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

const char *zo_cpu_isa(void) {
#if defined(__AVX512BW__)
    return "avx512bw";
#elif defined(__AVX2__)
    return "avx2";
#elif defined(__SSE2__)
    return "sse2";
#else
    return "generic";
#endif
}
}
