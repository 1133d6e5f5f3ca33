// sneaky_snake.cuh -- zoe's SneakySnake pre-alignment filter (src/alignment/sneaky_snake.rs:78-131) for a batch of
// (reference, query) pairs.  The outcome is Some(true) / Some(false) / None exactly as zoe decides it, including the f32
// threshold arithmetic.  Two kernels:
//   sneaky_snake_warp_kernel   one WARP per pair, both sequences staged in shared memory (coalesced loads).  The
//                              2E + 1 diagonals ("rows" of the chip maze) of one checkpoint are independent searches for
//                              the first obstacle, so lane = diagonal; the snake advances to the farthest obstacle
//                              (warp max) and succeeds as soon as any diagonal reaches the end condition (warp vote) --
//                              zoe's sequential row loop returns the same value because its early return does not
//                              depend on the order of the rows and last_col is a maximum.
//   sneaky_snake_kernel        one thread per pair with zoe's loops, for sequences beyond the staging area.
#pragma once
#include <cstdint>

namespace zoe_cuda {

struct SnakeParams {
    const uint8_t *refs;      // concatenated reference bytes
    const uint64_t *ref_off;  // [n + 1]
    const uint8_t *queries;   // concatenated query bytes
    const uint64_t *qry_off;  // [n + 1]
    uint64_t ref_base, qry_base;  // offset of the first byte held in refs / queries
    uint64_t n;
    float threshold;
    uint8_t *out;             // 0 = Some(false), 1 = Some(true), 2 = None
};

constexpr int kSnakeStage = 1024;  // bytes of staging per sequence per warp (longer pairs take the thread kernel)

// Some(..) / None decisions that need no maze: returns 0xff when the maze must be walked
__device__ __forceinline__ uint8_t snake_prologue(float threshold, uint64_t rl, uint64_t ql, uint64_t &edit_thresh,
                                                  uint64_t &len_diff) {
    if (!(threshold >= 0.0f && threshold <= 1.0f)) return 2;  // (0. ..=1.).contains(&threshold)
    edit_thresh = (uint64_t)floorf(__fmul_rn((float)ql, threshold));
    len_diff = rl > ql ? rl - ql : ql - rl;
    if (len_diff > edit_thresh) return 2;
    if (edit_thresh == ql) return 1;
    return 0xff;
}

__global__ void __launch_bounds__(256) sneaky_snake_warp_kernel(const SnakeParams p) {
    extern __shared__ __align__(16) uint8_t snake_sm[];
    constexpr unsigned ALL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    uint8_t *st1 = snake_sm + (size_t)warp * 2 * kSnakeStage, *st2 = st1 + kSnakeStage;
    for (uint64_t i = (uint64_t)blockIdx.x * wpb + warp; i < p.n; i += (uint64_t)gridDim.x * wpb) {
        const uint8_t *reference = p.refs + (p.ref_off[i] - p.ref_base), *query = p.queries + (p.qry_off[i] - p.qry_base);
        const uint64_t rl = p.ref_off[i + 1] - p.ref_off[i], ql = p.qry_off[i + 1] - p.qry_off[i];
        uint64_t edit_thresh = 0, len_diff = 0;
        uint8_t res = snake_prologue(p.threshold, rl, ql, edit_thresh, len_diff);
        if (res == 0xff) {
            const bool swap = rl > ql;  // choose the shorter string as s1
            const uint8_t *g1 = swap ? query : reference, *g2 = swap ? reference : query;
            const int n1 = (int)(swap ? ql : rl), n2 = (int)(swap ? rl : ql);
            const int E = (int)edit_thresh;
            __syncwarp();
            for (int k = lane; k < n1; k += 32) st1[k] = g1[k];
            for (int k = lane; k < n2; k += 32) st2[k] = g2[k];
            __syncwarp();
            const int window = 2 * E + 1, diffpad = (int)(len_diff / 2);
            int obstacles = 0, checkpoint = 0;
            bool passed = false;
            while (!passed && checkpoint < n1 && obstacles <= E && n1 - checkpoint > E - obstacles) {
                int last_col = checkpoint;
                for (int row0 = 0; row0 < window && !passed; row0 += 32) {
                    const int row = row0 + lane;
                    int mis = checkpoint;  // a diagonal without an obstacle leaves last_col alone
                    bool ok = false;
                    if (row < window) {
                        const int shift = row + diffpad - E;  // s2 index = col + shift
                        for (int col = checkpoint; col < n1; ++col) {
                            const int j = col + shift;
                            if (j >= 0 && j < n2 && st2[j] == st1[col]) {
                                if (col == n1 - 1 || n1 - col - 1 <= E - obstacles) {
                                    ok = true;
                                    break;
                                }
                            } else {
                                mis = col;
                                break;
                            }
                        }
                    }
                    passed = __any_sync(ALL, ok);
                    last_col = max(last_col, (int)__reduce_max_sync(ALL, (unsigned)mis));
                }
                checkpoint = last_col + 1;
                obstacles += 1;
            }
            res = passed ? 1 : (obstacles <= E ? 1 : 0);
        }
        if (lane == 0) p.out[i] = res;
    }
}

__global__ void __launch_bounds__(128) sneaky_snake_kernel(const SnakeParams p) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint8_t *reference = p.refs + (p.ref_off[i] - p.ref_base), *query = p.queries + (p.qry_off[i] - p.qry_base);
    const uint64_t rl = p.ref_off[i + 1] - p.ref_off[i], ql = p.qry_off[i + 1] - p.qry_off[i];
    uint8_t res;
    do {
        if (!(p.threshold >= 0.0f && p.threshold <= 1.0f)) {  // (0. ..=1.).contains(&threshold)
            res = 2;
            break;
        }
        const uint64_t edit_thresh = (uint64_t)floorf(__fmul_rn((float)ql, p.threshold));
        const uint64_t len_diff = rl > ql ? rl - ql : ql - rl;
        if (len_diff > edit_thresh) {
            res = 2;
            break;
        }
        if (edit_thresh == ql) {
            res = 1;
            break;
        }
        const bool swap = rl > ql;  // choose the shorter string as s1
        const uint8_t *s1 = swap ? query : reference, *s2 = swap ? reference : query;
        const uint64_t n1 = swap ? ql : rl, n2 = swap ? rl : ql;
        const uint64_t window = 2 * edit_thresh + 1, diffpad_len = len_diff / 2;
        uint64_t obstacles = 0, checkpoint = 0;
        bool passed = false;
        while (!passed && checkpoint < n1 && obstacles <= edit_thresh && n1 - checkpoint > edit_thresh - obstacles) {
            uint64_t last_col = checkpoint;
            for (uint64_t row = 0; row < window && !passed; ++row) {
                for (uint64_t col = checkpoint; col < n1; ++col) {
                    const uint64_t shifted = col + row + diffpad_len;
                    if (shifted >= edit_thresh && shifted - edit_thresh < n2 && s2[shifted - edit_thresh] == s1[col]) {
                        if (col == n1 - 1 || n1 - col - 1 <= edit_thresh - obstacles) {
                            passed = true;
                            break;
                        }
                    } else {
                        last_col = last_col > col ? last_col : col;
                        break;
                    }
                }
            }
            checkpoint = last_col + 1;
            obstacles += 1;
        }
        res = passed ? 1 : (obstacles <= edit_thresh ? 1 : 0);
    } while (false);
    p.out[i] = res;
}

}  // namespace zoe_cuda
