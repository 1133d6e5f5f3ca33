// sneaky_snake.cuh -- zoe's SneakySnake pre-alignment filter (src/alignment/sneaky_snake.rs:78-131) for a batch of
// (reference, query) pairs: one thread walks one pair's chip maze with zoe's loops (byte compares only; the outcome is
// Some(true) / Some(false) / None exactly as zoe decides it, including the f32 threshold arithmetic).
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
