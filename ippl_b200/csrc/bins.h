// bins.h -- host-side state of the cell-ordered particle store (ipplb_bins, include/ippl_b200.h).
#pragma once
#include "common.cuh"

namespace ipplb {

// per-buffer state words (device): where the unsorted tail starts and how many particles it holds
enum { BS_TAIL_START = 0, BS_TAIL_COUNT = 1, BS_WORDS = 8 };
constexpr int MAX_RANKS = 64;  // ranks of one NVSwitch domain the exit buffer is segmented for
// misc words (device): step scratch [0..3] (zeroed before every step) and the status block the plan kernel writes
enum {
    BM_WORK = 0, BM_EXIT = 1, BM_FLAGS = 2, BM_SPARE = 3,
    BM_ST_TOTAL = 8, BM_ST_TAIL = 9, BM_ST_EXIT = 10, BM_ST_FLAGS = 11, BM_ST_BUCKETED = 12, BM_ST_TAIL_START = 13,
    BM_WORDS = 16
};

}  // namespace ipplb

struct ipplb_bins {
    ipplb_mesh mesh;
    int ntx = 0, nty = 0, ntz = 0, ntiles = 0;
    long ncells   = 0;  // ntiles * 64 (tile-major key space)
    long capacity = 0;
    int cur       = 0;  // which table set describes the current particles
    bool built    = false;
    int* d_tab    = nullptr;  // start[2][nt] cap[2][nt] count[2][nt] state[2][BS_WORDS] misc[BM_WORDS]
    int* d_cell   = nullptr;  // build scratch: per-cell offsets [ncells + 1]
    long long* d_plan = nullptr;  // planning kernel scratch (partial sums + grid barrier words)
    int* d_exit_cnt = nullptr;    // [2][MAX_RANKS] leavers per destination rank: live counters, snapshot of the last step
    int exit_ranks  = 1;          // how many of them the last step used
    int exit_p2p    = 0;          // the last step wrote its leavers into the peers' inboxes (not into an exit buffer)
    // optional CUDA-event timing of the fused kernel alone, on the launching stream (ipplb_bins_set_timing)
    static constexpr int NEV = 256;
    cudaEvent_t* ev = nullptr;  // [2 * NEV]: start / stop per launch
    int ev_n = 0, timing = 0;
    int* h_status = nullptr;  // pinned [BM_WORDS]
    int build_variant = 1;    // ipplb_bins_set_build_variant: 1 per-cell positions, 2 arrival order (bins.cu)
    // slack = total / slack_div + slack_sqrt * sqrt(total) + slack_const  (elements per bucket)
    int slack_div = 32, slack_sqrt = 4, slack_const = 16;

    int* start(int b) const { return d_tab + (size_t)b * ntiles; }
    int* cap(int b) const { return d_tab + (size_t)(2 + b) * ntiles; }
    int* count(int b) const { return d_tab + (size_t)(4 + b) * ntiles; }
    int* state(int b) const { return d_tab + (size_t)6 * ntiles + b * ipplb::BS_WORDS; }
    int* misc() const { return d_tab + (size_t)6 * ntiles + 2 * ipplb::BS_WORDS; }
    size_t tab_words() const { return (size_t)6 * ntiles + 2 * ipplb::BS_WORDS + ipplb::BM_WORDS; }
};

namespace ipplb {
// clamps count[o] to cap[o], plans start/cap of the other buffer from the totals, zeroes its cursors, writes
// the status block.  Enqueued on the context's stream.
int bins_plan(ipplb_ctx* ctx, ipplb_bins* b, int o, int seg_cap = 0x7fffffff);
}  // namespace ipplb
