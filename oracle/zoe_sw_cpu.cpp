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

/* shift_elements_right::<1>(fill) */
template <typename T, int N>
static inline V<T, N> shr1(V<T, N> v, T fill) {
    alignas(64) T tmp[N + 1];
    tmp[0] = fill;
    std::memcpy(tmp + 1, &v, sizeof(T) * (N - 1));
    V<T, N> out;
    std::memcpy(&out, tmp, sizeof(out));
    return out;
}

template <typename T, int N>
static inline bool any_gt(V<T, N> a, V<T, N> b) {
    auto m = a > b;
    constexpr int W = N * sizeof(T) / 8;
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
                         std::vector<V<T, N>> &scratch, uint32_t *score) {
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
    T best = max_scores[0];
    for (int i = 1; i < N; i++) best = std::max(best, max_scores[i]);
    if (!(best < MAX)) return 1;
    uint32_t s = (uint32_t)((int64_t)MAX + 1 + (int64_t)best);
    *score = s;
    return s == 0 ? 2 : 0;
}

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
        for (uint64_t i = a; i < b; i++) {
            const uint8_t *r = reads + off[i];
            size_t len = (size_t)(off[i + 1] - off[i]);
            for (uint32_t j = 0; j < n_prof; j++) {
                ProfileSet<M, N, O> &ps = *sets[j];
                uint32_t sco = 0;
                int t = 8;
                int rc = striped_score<int8_t, M>(ps.p8, r, len, sc.map, s8, &sco);
                if (rc == 1) {  // or_else_overflowed: output.rs:81-83
                    std::call_once(ps.f16, [&] { ps.p16.build(ps.seq, ps.m, sc); });
                    t = 16;
                    rc = striped_score<int16_t, N>(ps.p16, r, len, sc.map, s16, &sco);
                    if (rc == 1) {
                        std::call_once(ps.f32, [&] { ps.p32.build(ps.seq, ps.m, sc); });
                        t = 32;
                        rc = striped_score<int32_t, O>(ps.p32, r, len, sc.map, s32, &sco);
                    }
                }
                size_t k = (size_t)i * n_prof + j;
                score[k] = rc == 0 ? sco : 0;
                status[k] = (uint8_t)rc;
                tier[k] = (uint8_t)t;
            }
        }
    };
    if (n_threads <= 1) {
        worker(0, n);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(worker, n * t / n_threads, n * (t + 1) / n_threads);
    for (auto &x : th) x.join();
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
