// sched.cpp — host-side work placement for the multi-GPU depth path.
// The reference's unit of independence is the 10 Mb chunk / the reference sequence (depth/depth.go:129-159: one
// `samtools depth` child per chunk, GOMAXPROCS of them at a time); here whole contigs are dealt to the GPUs of one box,
// longest first onto the least loaded GPU (LPT), no data-path collective.  The CLI (`goleft depth --gpus N`) and bench.py
// use the same function, so what is benchmarked is what ships.
#include <stdint.h>
#include <algorithm>
#include <numeric>
#include <vector>
#include "../../../include/goleft_b200.h"

extern "C" int gl_lpt_assign(const int64_t* weight, int32_t n, int32_t bins, int32_t* bin_of, int64_t* bin_load) {
    if (n < 0 || bins <= 0 || (n > 0 && (!weight || !bin_of))) return GL_EINVAL;
    std::vector<int32_t> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return weight[a] > weight[b]; });
    std::vector<int64_t> load((size_t)bins, 0);
    for (int32_t i : order) {
        const int32_t b = (int32_t)(std::min_element(load.begin(), load.end()) - load.begin());   // ties: lowest bin
        bin_of[i] = b;
        load[(size_t)b] += weight[i] > 0 ? weight[i] : 0;
    }
    if (bin_load) for (int32_t b = 0; b < bins; b++) bin_load[b] = load[(size_t)b];
    return GL_OK;
}
