// tc_scan_wide.cuh — kernel W: the posting-list scan with WIDE query tiles (256 queries per work
// item) on CTA pairs.  Included by tc_scan.cu inside its anonymous namespace (shares TcScanParams,
// ScanItem and the stopwatch macros with the other kernels).
//
// Replaces the row loop of IVFIndex::search_with_config (src/ivf/core.rs:661-674) for a whole batch,
// and — with identity items — the exhaustive scan that stands in for HNSWIndex::search
// (src/hnsw/core.rs:398-467) and the argmin of find_nearest_centroid (src/ivf/core.rs:373-386).
//
// Why.  Kernel R keeps the query tile of an item in shared memory (64 queries x 1536 B = 96 KB), so
// a list probed by more than 64 queries is streamed through the SMs once per 64-query group: 2.06 M
// rows entered the SMs for 1.0 M distinct rows on the bench index.  Here the queries are the MMA's A
// operand and live in TENSOR MEMORY: 128 queries per CTA (TMEM lane = query, D <= 384 columns), 256
// per CTA pair (tcgen05 cta_group::2, M = 256).  The database rows are the B operand, and the pair
// SPLITS it: of every 64-row tile each CTA streams 32 rows (TMA, 128-byte swizzle) into its own
// shared-memory ring, so a row enters exactly one SM once per 256-query item.  Shared memory holds
// nothing but the ring (8 x 16 KB in flight per SM) and the per-query candidate heaps.
//
// Epilogue: one thread owns one (query, half tile): accumulator lane x 32 of the tile's 64 columns.
// v = |x|^2 - 2 q.x against the thread's threshold in a register; a passing row goes into the
// thread's own 32-entry max-heap (shared memory, one row per thread: no atomics, no votes, no
// warp-wide merges — 32 lanes insert into 32 heaps in the same instructions).  The heap's root is a
// running bound of the query's 32nd-best distance; it is shared with every other CTA and warp that
// scans rows for the same query through thr_g (and, multi-GPU, pushed to the peer GPUs).  At the
// end of an item the heaps are published as they are (unsorted rows of 32 keys; the shortlist merge
// after the scan sorts them): TWO shortlist rows per (query, item), one per half.
//
// Warp roles per CTA (512 threads, four warpgroups; registers are moved from the first to the last
// with setmaxnreg):
//   warp 0        TMA producer (+ the dynamic tile scheduler in the leader)
//   warp 1        MMA issuer (leader only), TMEM owner
//   warp 2        |x|^2 strips (per tile: the row norms, +inf for rows that are tombstoned, filtered
//                 out or past the end of the item; four tiles of loads in flight, so the DRAM latency
//                 of the norm array never sits on the TMA issue path)
//   warps 4-7     epilogue, columns 0-31 of every tile   (warp & 3 = TMEM lane quarter)
//   warps 8-11    epilogue, columns 32-63
//   warps 12-15   query loaders: global -> shared-memory transposition -> tcgen05.st into the A
//                 region; the loads of the next item's first four k-blocks are in flight while the
//                 current item is still being multiplied
//
// Barriers that another CTA arrives on use the default CTA-scope semantics, as CUTLASS's
// ClusterBarrier does: what they order is tcgen05 / TMA traffic, which carries its own fences and
// transaction counts.  (A .release.cluster arrival costs MEMBAR.ALL.GPU, an .acquire.cluster wait
// CCTL.IVALL — measured: 1100 cycles per tile.)  Only the work-item ring, where the leader writes
// the peer's shared memory with a plain store, keeps the cluster-scope pair.
#pragma once

constexpr int W_M = 128;                         // queries per CTA (TMEM lanes)
constexpr int W_NQ = 2 * W_M;                    // queries per work item (the pair's MMA M)
constexpr int W_N = 64;                          // rows per tile of the pair (MMA N)
constexpr int W_NH = W_N / 2;                    // ... of which each CTA streams half
constexpr int W_KBLK_BYTES = W_NH * 128;         // 4 KB: 32 rows x one 128-byte k-block
constexpr int W_ACC_COL = 384;                   // accumulators: tensor-memory columns 384..511
constexpr int W_NBUF = 2;                        // accumulator buffers of W_N columns
constexpr int W_NSLOT = 8;                       // |x|^2 strips in flight
constexpr int W_HEAP_BYTES = 34 * 8;             // one heap: entries 1..32 of {row, d2 bits} (1-based: the children of i are the
                                                 // 16-byte pair at 16 i), entry 33 a zero sentinel; 272 B = 17 x 16 B keeps the
                                                 // lanes of a quarter warp on distinct bank groups
constexpr int W_HEAPS = 2 * W_M;                 // heaps per CTA: (column half, query)
constexpr int W_XP_LD = 32;                      // floats per query row of a loader warp's transposition buffer (16-byte chunks XOR-swizzled)
constexpr int W_PEND = 4;                        // candidates a thread parks before it folds them into its heap
constexpr int W_PEND_BYTES = (W_PEND + 1) * 8;   // parked {row, d2 bits} pairs of one thread (+ 8 B: odd stride in 8-byte units)
constexpr int W_FLUSH_TILES = 4;                 // every warp folds its parked candidates once per this many tiles
constexpr int W_THREADS = 512;
constexpr int W_TMEM_COLS = 512;
constexpr int W_SUBROWS = 2;                     // shortlist rows per (query, item): one per column half
static_assert(W_ACC_COL + W_NBUF * W_N <= W_TMEM_COLS, "tensor memory budget of kernel W");
constexpr int W_NBARS_FIXED = 2 * W_NBUF + 2 * W_NSLOT + 2 * TC_SCHED + 2;   // + 2 * STAGES

size_t tc_scan_wide_smem_bytes(uint32_t stages, uint32_t kbs) {
    return (size_t)stages * kbs * W_KBLK_BYTES + (size_t)W_HEAPS * W_HEAP_BYTES + (size_t)4 * 32 * W_XP_LD * 4 +
           (size_t)W_HEAPS * W_PEND_BYTES + (size_t)W_NSLOT * W_N * 4 + (size_t)(2 * stages + W_NBARS_FIXED) * 8 + 16 + (size_t)TC_SCHED * 4;
}
uint32_t wide_pick_kbs(uint32_t KB) {
    for (uint32_t c : {4u, 3u, 2u}) if (KB % c == 0) return c;
    return 1;
}
// deepest ring that fits shared memory
uint32_t wide_pick_stages(uint32_t KB, uint32_t kbs) {
    (void)KB;
    uint32_t stages = 12u;
    while (stages > 2 && tc_scan_wide_smem_bytes(stages, kbs) + 1024 > 232448) --stages;
    return stages;
}

// ---- thread-private candidate heaps (shared memory, addressed with explicit ld/st.shared) --------
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_v2(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}

// A thread's candidate list (shared address hb; entry i = {row, d2 bits} at hb + 8 i, i = 1..32) has
// two phases.  FILLING: fewer than 32 candidates have passed the shared bound so far; they are stored
// in arrival order, nothing else (most (query, item) pairs never leave this phase: what they hold is
// published as it is).  HEAP: at the 32nd candidate the entries are arranged as a max-heap once
// (wide_heapify); from then on the root is the thread's own bound, and a candidate below it replaces
// it (wide_fold, sift down: the two children of entry i are ONE 16-byte load at hb + 16 i).

// sift the pair (row, d) down from entry i of a full heap
__device__ __forceinline__ void wide_sift_down(uint32_t hb, uint32_t i, uint32_t row, uint32_t d) {
    while (i <= (uint32_t)TC_KP / 2) {
        const uint4 ch = lds_v4(hb + 16 * i);   // children 2 i and 2 i + 1 (entry 33 is a zero sentinel)
        const bool right = ch.w > ch.y;
        const uint32_t cd = right ? ch.w : ch.y, cp = right ? ch.z : ch.x;
        if (cd <= d) break;
        sts_v2(hb + 8 * i, cp, cd);
        i = 2 * i + (right ? 1u : 0u);
    }
    sts_v2(hb + 8 * i, row, d);
}
// Floyd's heap construction over the 32 stored entries (lanes whose list just filled: `mine`); returns the root
__device__ __noinline__ uint32_t wide_heapify(uint32_t hb, bool mine) {
    if (!mine) return 0xFFFFFFFFu;
    for (uint32_t i = (uint32_t)TC_KP / 2; i >= 1; --i) {
        const uint2 e = lds_v2(hb + 8 * i);
        wide_sift_down(hb, i, e.x, e.y);
    }
    return lds_v2(hb + 8).y;
}
// Fold a thread's `pcnt` parked candidates (shared address pb) into its full heap: each one below the
// root replaces it.  Returns the new root.  All lanes of a warp call this together: the warp pays
// max-over-lanes(parked) insertions, not their sum.
__device__ __noinline__ uint32_t wide_fold(uint32_t hb, uint32_t pb, uint32_t pcnt, uint32_t root) {
    for (uint32_t e = 0; e < pcnt; ++e) {
        const uint2 c = lds_v2(pb + 8 * e);          // {row, d2 bits}
        if (c.y < root) {
            wide_sift_down(hb, 1, c.x, c.y);
            root = lds_v2(hb + 8).y;
        }
    }
    return root;
}

// METRIC (flat-tier scans of a similarity handle, include/fvdb.h fvdb_metric): the accumulator is q.x either
// way; what changes is the per-row strip and the map from accumulator to the non-negative, smaller-is-better
// approximate key the heaps, bounds and the shortlist merge work on:
//   L2   strip |x|^2,       v = strip - 2 acc,   key = v + |q|^2                  (approximate d^2)
//   COS  strip -1 / |x|,    v = strip * acc,     key = v / |q| + 1                (1 - cosine; a zero row scores 0)
//   DOT  strip -1,          v = strip * acc,     key = v + B, B = 1.02 |q| max|x| (B - dot >= 0)
// Masked rows carry +inf (L2) / NaN (COS, DOT) in the strip and never pass `v < threshold`.
// SMALL (nearest-centroid assignment: k = 1, four candidates re-ranked, the fifth is the proof bound): a thread keeps
// its 8 best candidates in REGISTERS, sorted, instead of the 32-entry heap in shared memory — a running top-8 of
// the 2048 centroids a thread sees takes ~45 insertions of eight min / max pairs instead of ~165 heap updates, and
// its bound tightens four times faster.  At the end of the item the list is spilled into the heap's slots and
// published through the common path (rows of 32 keys, 8 of them valid).
constexpr int W_SMALL_N = 8;
template <int METRIC, bool SMALL = false>
__global__ void __launch_bounds__(W_THREADS, 1)
tc_scan_wide_kernel(const __grid_constant__ CUtensorMap tmap, const TcScanParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t KB = p.KB;
    const uint32_t KBS = p.kbs;                  // k-blocks per ring stage
    const uint32_t NST = KB / KBS;               // stages per row tile
    const uint32_t STAGES = p.stages;
    const uint32_t STAGE_BYTES = KBS * W_KBLK_BYTES;
    unsigned char* ring = smem;                                                       // STAGES x KBS x 4 KB
    unsigned char* heaps = ring + (size_t)STAGES * STAGE_BYTES;                       // [W_HEAPS] x 272 B max-heaps of {row, d2 bits}
    unsigned char* parks = heaps + (size_t)W_HEAPS * W_HEAP_BYTES;                    // [W_HEAPS] x 40 B parked candidates
    float* xp = reinterpret_cast<float*>(parks + (size_t)W_HEAPS * W_PEND_BYTES);     // [4][32][32] loader transposition
    float* xn_ring = xp + 4 * 32 * W_XP_LD;                                           // [W_NSLOT][W_N]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xn_ring + W_NSLOT * W_N);
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + W_NBARS_FIXED);
    uint32_t* sched_s = tmem_ptr_s + 1;

    const uint32_t rank = cluster_ctarank();     // 0 = leader
    const bool leader = rank == 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = smem_u32(bars);                   // leader only: both halves of a stage have landed
    const uint32_t bar_empty = bar_full + 8 * STAGES;           // the MMAs reading a stage have retired (multicast)
    const uint32_t bar_tfull = bar_empty + 8 * STAGES;          // accumulator ready (multicast)
    const uint32_t bar_tempty = bar_tfull + 8 * W_NBUF;         // leader only: both CTAs' epilogues released it
    const uint32_t bar_nfull = bar_tempty + 8 * W_NBUF;         // norm strip of a tile staged (own CTA)
    const uint32_t bar_nempty = bar_nfull + 8 * W_NSLOT;        // ... and read by the eight epilogue warps
    const uint32_t bar_sfull = bar_nempty + 8 * W_NSLOT;        // work item published
    const uint32_t bar_sempty = bar_sfull + 8 * TC_SCHED;       // leader only: every consumer of both CTAs took it
    const uint32_t bar_qready = bar_sempty + 8 * TC_SCHED;      // leader only: both CTAs' queries sit in tensor memory
    const uint32_t bar_qfree = bar_qready + 8;                  // every MMA of the item has retired (multicast)
    const uint32_t l_full = mapa_u32(bar_full, 0), l_tempty = mapa_u32(bar_tempty, 0);
    const uint32_t l_sempty = mapa_u32(bar_sempty, 0), l_qready = mapa_u32(bar_qready, 0);

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int i = 0; i < W_NBUF; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 16);   // eight epilogue warps of each CTA
        }
        for (int i = 0; i < W_NSLOT; ++i) {
            mbar_init(bar_nfull + 8 * i, 1);
            mbar_init(bar_nempty + 8 * i, 8);
        }
        for (int i = 0; i < TC_SCHED; ++i) {
            mbar_init(bar_sfull + 8 * i, 1);
            // leader: MMA + strip + 8 epilogue + 4 loader warps; peer: producer + strip + 8 epilogue + 4 loader warps
            mbar_init(bar_sempty + 8 * i, 28);
        }
        mbar_init(bar_qready, 8);                // four loader warps of each CTA
        mbar_init(bar_qfree, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < W_HEAPS) sts_v2(smem_u32(heaps) + threadIdx.x * W_HEAP_BYTES + 8 * 33, 0u, 0u);   // sentinel child of entry 16
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                     "r"((uint32_t)W_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // barriers of both CTAs initialised, tensor memory of both allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;
    const uint32_t n_items = *p.item_count;
    unsigned long long lap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tl = clock64();
    if (p.prof && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.prof[((size_t)blockIdx.x * 6 + 3) * 8 + 4] = gt;
        p.prof[((size_t)blockIdx.x * 6 + 3) * 8 + 6] = (unsigned long long)tl;
    }
    // consumer side of the item ring: every consumer of both CTAs releases a slot on the LEADER's barrier
    auto release_slot = [&](uint32_t ss) {
        if (leader) mbar_arrive(bar_sempty + 8 * ss);
        else mbar_arrive_remote(l_sempty + 8 * ss);
    };
    // next work item of a consumer role (the peer's copy of the ring is written by the leader over
    // distributed shared memory: cluster-scope acquire)
    auto next_item = [&](uint32_t& ss, uint32_t& sphase) -> uint32_t {
        if (leader) mbar_wait(bar_sfull + 8 * ss, sphase);
        else mbar_wait_cl(bar_sfull + 8 * ss, sphase);
        const uint32_t item = *reinterpret_cast<volatile uint32_t*>(sched_s + ss);
        __syncwarp();
        if (lane == 0) release_slot(ss);
        if (++ss == TC_SCHED) { ss = 0; sphase ^= 1; }
        return item;
    };
    // the item query a thread owns in the epilogue / loader roles: queries alternate between the two
    // CTAs and are dealt round-robin to the four lane quarters, so that an item with few queries
    // still spreads its candidate handling over both CTAs and all epilogue warps
    const uint32_t quarter = (uint32_t)warp & 3u;
    const uint32_t jq = 2u * (4u * (uint32_t)lane + quarter) + rank;

    if (warp < 4) {
        // ======================= warpgroup 0: producer, MMA issuer, strips =======================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 0) {
            // ============ TMA producer (both CTAs) + tile scheduler (leader) ============
            uint32_t stage = 0, phase = 0, ss = 0, sphase = 0;
            const uint64_t hint_first = 0x12F0000000000000ull;   // L2 evict-first: rows read once
            const uint64_t hint_normal = 0x1000000000000000ull;  // rows shared by several items
            const uint32_t ring_base = smem_u32(ring);
            const uint32_t peer_sched = mapa_u32(smem_u32(sched_s), 1), peer_sfull = mapa_u32(bar_sfull, 1);
            while (true) {
                uint32_t item = 0;
                if (leader) {
                    Q1_LAP(3);
                    mbar_wait(bar_sempty + 8 * ss, sphase ^ 1);
                    if (lane == 0) {
                        item = atomicAdd(p.work_counter, 1u);
                        if (item >= n_items) item = ITEM_END;
                        sched_s[ss] = item;
                        st_cluster_u32(peer_sched + 4 * ss, item);
                        mbar_arrive(bar_sfull + 8 * ss);
                        mbar_arrive_cluster(peer_sfull + 8 * ss);
                    }
                    item = __shfl_sync(0xffffffffu, item, 0);
                    if (++ss == TC_SCHED) { ss = 0; sphase ^= 1; }
                } else {
                    item = next_item(ss, sphase);
                }
                Q1_LAP(0);
                if (item == ITEM_END) break;
                const ScanItem it = p.items[item];
                if (it.pair_count == 0 || it.row_begin >= it.row_end) continue;
                lap[6] += 1;
                Q1_LAP(1);
                const uint64_t hint = (it.identity == 2 || (!it.identity && it.slot > 1)) ? hint_normal : hint_first;
                // tiles of 64 rows: this CTA streams rows [rt + 32 * rank, + 32) of each
                for (uint32_t rt = it.row_begin; rt < it.row_end; rt += W_N) {
                    const uint32_t my_rt = rt + rank * W_NH;
                    const bool mine_live = my_rt < it.row_end;
                    const bool peer_live = rt + W_NH < it.row_end;   // rank 1's half holds rows of this item
                    lap[7] += 1;
                    for (uint32_t st = 0; st < NST; ++st) {
                        Q1_LAP(3);
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        Q1_LAP(2);
                        if (elect_one()) {
                            // the leader arms its barrier with the bytes of both CTAs' boxes; a half tile past
                            // the end of the item is not loaded (its rows are masked by +inf norms)
                            if (leader) mbar_expect_tx(bar_full + 8 * stage, STAGE_BYTES * (peer_live ? 2u : 1u));
                            if (mine_live) {
                                const uint32_t dst = ring_base + stage * STAGE_BYTES;
                                for (uint32_t j = 0; j < KBS; ++j)
                                    tma_load_2d_pair(dst + j * W_KBLK_BYTES, &tmap, l_full + 8 * stage,
                                                     (int)((st * KBS + j) * TC_KB_FLOATS), (int)my_rt, hint);
                            }
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
            Q1_LAP_DUMP(0);
        } else if (warp == 1) {
            // ============ MMA issuer: the leader's warp only (M = 256 over the pair, N = 64 rows) ============
            if (leader) {
                uint32_t stage = 0, phase = 0, ss = 0, sphase = 0, tile = 0, nit = 0;
                const uint32_t ring_base = smem_u32(ring);
                // timing experiments (FVDB_TC_DEBUG): bit 5 = N = 32 per instruction, bit 4 = a quarter of the instructions
                const uint32_t idesc = umma_idesc_tf32(2 * W_M, (p.debug & 32u) ? W_N / 2 : W_N);
                const uint32_t k4_n = (p.debug & 16u) ? 1u : 4u;
                while (true) {
                    Q1_LAP(4);
                    const uint32_t item = next_item(ss, sphase);
                    Q1_LAP(0);
                    if (item == ITEM_END) break;
                    const ScanItem it = p.items[item];
                    if (it.pair_count == 0 || it.row_begin >= it.row_end) continue;
                    Q1_LAP(4);
                    mbar_wait(bar_qready, nit & 1u);
                    tc_fence_after();
                    Q1_LAP(1);
                    for (uint32_t rt = it.row_begin; rt < it.row_end; rt += W_N) {
                        const uint32_t buf = tile & (W_NBUF - 1);
                        Q1_LAP(4);
                        mbar_wait(bar_tempty + 8 * buf, ((tile / W_NBUF) & 1u) ^ 1u);
                        tc_fence_after();
                        Q1_LAP(2);
                        const uint32_t d_tmem = tmem_base + W_ACC_COL + buf * W_N;
                        const bool last_tile = rt + W_N >= it.row_end;
                        for (uint32_t st = 0; st < NST; ++st) {
                            Q1_LAP(4);
                            mbar_wait(bar_full + 8 * stage, phase);
                            tc_fence_after();
                            Q1_LAP(3);
                            if (elect_one()) {
                                const uint32_t sbase = ring_base + stage * STAGE_BYTES;
                                if (!(p.debug & 8u)) {
                                    for (uint32_t j = 0; j < KBS; ++j) {
                                        const uint64_t b0 = umma_desc_sw128(sbase + j * W_KBLK_BYTES);
                                        const uint32_t a0 = tmem_base + (st * KBS + j) * 32;
#pragma unroll
                                        for (uint32_t k4 = 0; k4 < 4; ++k4)   // A: 8 tf32 = 8 TMEM columns per step
                                            if (k4 < k4_n) umma_tf32_ts_pair(d_tmem, a0 + k4 * 8, b0 + 2 * k4, idesc, (st | j | k4) != 0 ? 1u : 0u);
                                    }
                                }
                                umma_commit_pair(bar_empty + 8 * stage);    // frees the stage in both CTAs
                                if (st + 1 == NST) {
                                    umma_commit_pair(bar_tfull + 8 * buf);  // accumulators ready in both CTAs
                                    if (last_tile) umma_commit_pair(bar_qfree);  // the query tiles may be replaced
                                }
                            }
                            __syncwarp();
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                        ++tile;
                    }
                    ++nit;
                }
                Q1_LAP_DUMP(1);
            }
        } else if (warp == 2) {
            // ============ |x|^2 strips: four tiles per round, their loads in flight together ============
            uint32_t ss = 0, sphase = 0, tcount = 0;
            while (true) {
                const uint32_t item = next_item(ss, sphase);
                if (item == ITEM_END) break;
                const ScanItem it = p.items[item];
                if (it.pair_count == 0 || it.row_begin >= it.row_end) continue;
                for (uint32_t rt = it.row_begin; rt < it.row_end; rt += 4 * W_N) {
                    float xnv[8];
                    uint32_t idv[8];
#pragma unroll
                    for (int h = 0; h < 8; ++h) {   // tile h / 2, rows (h & 1) * 32 + lane of it
                        const uint32_t pos = rt + h * 32 + lane;
                        xnv[h] = __uint_as_float(F32_INF_BITS);
                        idv[h] = pos;
                        if (pos < it.row_end) {
                            xnv[h] = __ldg(p.xnorm + pos);
                            if ((p.tomb || p.filt) && p.ids) idv[h] = __ldg(p.ids + pos);
                        }
                    }
                    if (p.tomb || p.filt) {
#pragma unroll
                        for (int h = 0; h < 8; ++h) {
                            const uint32_t pos = rt + h * 32 + lane;
                            if (pos < it.row_end) {
                                bool live = true;
                                if (p.tomb && bit_test(p.tomb, p.tomb_bits, idv[h])) live = false;
                                else if (p.filt && !bit_test(p.filt, p.filt_bits, idv[h])) live = false;
                                if (!live) xnv[h] = __uint_as_float(F32_INF_BITS);
                            }
                        }
                    }
                    if (METRIC != METRIC_L2) {
#pragma unroll
                        for (int h = 0; h < 8; ++h) {
                            const float xn = xnv[h];
                            if (xn == __uint_as_float(F32_INF_BITS)) xnv[h] = __uint_as_float(0x7fc00000u);   // masked: NaN
                            else if (METRIC == METRIC_DOT) xnv[h] = -1.0f;
                            else xnv[h] = xn > 0.0f ? -rsqrtf(xn) : 0.0f;
                        }
                    }
#pragma unroll
                    for (int t4 = 0; t4 < 4; ++t4) {
                        if (rt + t4 * W_N >= it.row_end) break;
                        const uint32_t slot = tcount % W_NSLOT;
                        mbar_wait(bar_nempty + 8 * slot, ((tcount / W_NSLOT) & 1u) ^ 1u);
                        xn_ring[slot * W_N + lane] = xnv[2 * t4];
                        xn_ring[slot * W_N + 32 + lane] = xnv[2 * t4 + 1];
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_nfull + 8 * slot);
                        ++tcount;
                    }
                }
            }
        }
    } else if (warp >= 12) {
        // ============ warpgroup 3: query loaders, this CTA's 128 queries -> tensor memory (A operand) ============
        // Rows are read coalesced (8 lanes x 16 B = one 128-byte k-block of one query, four queries per
        // load instruction), transposed through this warp's shared-memory buffer, and written by the
        // owning lane (TMEM lane = query) with one 32-column tcgen05.st per k-block.  Four k-blocks of
        // loads are in flight; those of the first four are issued BEFORE the wait for the previous
        // item's MMAs.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const uint32_t D = p.D;
        const uint32_t lane_taddr = (quarter * 32u) << 16;
        float* xpw = xp + quarter * 32 * W_XP_LD;
        uint32_t ss = 0, sphase = 0, nit = 0;
        while (true) {
            Q1_LAP(3);
            const uint32_t item = next_item(ss, sphase);
            Q1_LAP(0);
            if (item == ITEM_END) break;
            const ScanItem it = p.items[item];
            if (it.pair_count == 0 || it.row_begin >= it.row_end) continue;
            uint32_t qi = ID_NONE;
            if (jq < it.pair_count) qi = it.identity ? it.pair_begin + jq : p.pair_q[it.pair_begin + jq];
            const unsigned havem = __ballot_sync(0xffffffffu, qi != ID_NONE);
            const float4* src[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t qs = __shfl_sync(0xffffffffu, qi, 4 * i + (lane >> 3));
                src[i] = (qs == ID_NONE) ? nullptr : reinterpret_cast<const float4*>(p.Q + (size_t)qs * D) + (lane & 7);
            }
            auto ldq = [&](uint32_t kb, float4 (&v)[8]) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if ((havem >> (4 * i)) & 0xFu)      // warp-uniform: some query of this instruction exists
                        v[i] = src[i] ? __ldg(src[i] + kb * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            // one k-block: registers -> transposition buffer -> this lane's 32 columns of tensor memory;
            // the buffer's registers are refilled with k-block kb + 4 right away
            auto stage_kb = [&](uint32_t kb, float4 (&c)[8]) {
                if (kb >= KB) return;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<float4*>(xpw + (4 * i + (lane >> 3)) * W_XP_LD +
                                               (((lane & 7) ^ ((4 * i + (lane >> 3)) & 7)) << 2)) = c[i];
                if (kb + 4 < KB) ldq(kb + 4, c);
                __syncwarp();
                uint32_t r[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(xpw + lane * W_XP_LD + ((i ^ (lane & 7)) << 2));
                    r[4 * i + 0] = __float_as_uint(v.x); r[4 * i + 1] = __float_as_uint(v.y);
                    r[4 * i + 2] = __float_as_uint(v.z); r[4 * i + 3] = __float_as_uint(v.w);
                }
                __syncwarp();   // the buffer may be overwritten by the next k-block
                tmem_st32(tmem_base + lane_taddr + kb * 32, r);
            };
            float4 c0[8], c1[8], c2[8], c3[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) c0[i] = c1[i] = c2[i] = c3[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (havem) {
                ldq(0, c0);
                if (KB > 1) ldq(1, c1);
                if (KB > 2) ldq(2, c2);
                if (KB > 3) ldq(3, c3);
            }
            Q1_LAP(1);
            // the A region may be overwritten once every MMA of the previous item has retired
            if (nit >= 1) mbar_wait(bar_qfree, (nit - 1) & 1u);
            tc_fence_after();
            Q1_LAP(3);
            if (havem) {
                for (uint32_t kb = 0; kb < KB; kb += 4) {
                    stage_kb(kb, c0);
                    stage_kb(kb + 1, c1);
                    stage_kb(kb + 2, c2);
                    stage_kb(kb + 3, c3);
                }
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(bar_qready);
                else mbar_arrive_remote(l_qready);
            }
            ++nit;
            Q1_LAP(2);
        }
        if (warp == 12 && p.prof && lane == 0) { for (int i_ = 0; i_ < 4; ++i_) p.prof[((size_t)blockIdx.x * 6 + 3) * 8 + i_] = lap[i_]; }
    } else {
        // ======= warpgroups 1, 2: epilogue.  One thread = one query x one half (32 columns) of every tile =======
        const uint32_t half = ((uint32_t)warp - 4u) >> 2;    // 0: columns 0-31 (rows streamed by the leader), 1: 32-63
        const uint32_t m = quarter * 32u + (uint32_t)lane;   // TMEM lane == query row of this CTA's tile
        const uint32_t lane_taddr = (quarter * 32u) << 16;
        const uint32_t hb = smem_u32(heaps) + (half * W_M + m) * W_HEAP_BYTES;   // this thread's heap
        const uint32_t pb = smem_u32(parks) + (half * W_M + m) * W_PEND_BYTES;   // ... and its parked candidates
        uint32_t ss = 0, sphase = 0, tile = 0, tcount = 0;
        uint32_t st_app = 0, st_rounds = 0, st_chunks = 0;
        while (true) {
            Q1_LAP(0);
            const uint32_t item = next_item(ss, sphase);
            Q1_LAP(0);
            if (item == ITEM_END) break;
            const ScanItem it = p.items[item];
            if (it.pair_count == 0 || it.row_begin >= it.row_end) continue;
            // ---- item prologue: this thread's query, its threshold from the shared bound ----
            const bool have = jq < it.pair_count;
            uint32_t qi = 0, sl = 0;
            // key = v * qa + qn (see the table above the kernel); a bound in key units maps back to v units
            // as (bound - qn) * qia.  L2: qa = qia = 1 and qn = |q|^2, so both maps are the plain +- |q|^2.
            float qn = 0.f, qa = 1.0f, qia = 1.0f, thrp = -__uint_as_float(F32_INF_BITS);   // lanes without a query never pass
            uint32_t thr_pending = F32_INF_BITS, root_pub = F32_INF_BITS, peer_sent = F32_INF_BITS;
            if (have) {
                if (it.identity) { qi = it.pair_begin + jq; sl = it.slot; }
                else { qi = p.pair_q[it.pair_begin + jq]; sl = p.pair_slot[it.pair_begin + jq]; }
                qn = p.qnorm[qi];
                if (METRIC == METRIC_COS) {
                    qa = qn > 0.0f ? rsqrtf(qn) : 0.0f;     // a zero query scores 0 everywhere: nothing passes, and
                    qia = qn > 0.0f ? sqrtf(qn) : 0.0f;     // the re-rank hands it to the exact path
                    qn = 1.0f;
                } else if (METRIC == METRIC_DOT) {
                    qn = 1.02f * sqrtf(qn) * sqrtf(__uint_as_float(*p.xmax_bits));
                }
                if (p.thr_g) thr_pending = *(volatile uint32_t*)(p.thr_g + qi);
                thrp = (__uint_as_float(thr_pending) - qn) * qia;
            }
            unsigned long long best[W_SMALL_N];   // SMALL: this thread's candidates, ascending keys (approx d2 bits << 32 | row)
#pragma unroll
            for (int i = 0; i < W_SMALL_N; ++i) best[i] = KEY_NONE;
            uint32_t hcnt = 0;   // entries in this thread's candidate list (filling, then a max-heap on the approx d2 bits)
            uint32_t pcnt = 0;   // candidates parked since the last fold (heap phase)
            // In the heap phase a candidate costs a chain of dependent shared-memory accesses, and
            // candidates are rare (one per lane every few tiles), so inserting them as they come would
            // serialise the warp: one lane works, 31 wait.  They are parked and folded by all lanes together.
            uint32_t root = 0xFFFFFFFFu;   // the heap's root once the list is full (hcnt == 32)
            auto fold = [&]() {
                root = wide_fold(hb, pb, pcnt, root);
                thrp = fminf(thrp, (__uint_as_float(root) - qn) * qia);
                pcnt = 0;
            };
            Q1_LAP(0);

            // ---- row tiles ----
            for (uint32_t rt = it.row_begin; rt < it.row_end; rt += W_N) {
                // bound tightened meanwhile by the other half's warp and by CTAs scanning other lists of
                // the same query (loaded behind the release of an earlier accumulator: no exposed latency)
                if (have) thrp = fminf(thrp, (__uint_as_float(thr_pending) - qn) * qia);
                if (p.debug & 2u) thrp = -__uint_as_float(F32_INF_BITS);   // timing experiment: nothing ever passes
                const uint32_t slot = tcount % W_NSLOT;
                Q1_LAP(0);
                mbar_wait(bar_nfull + 8 * slot, (tcount / W_NSLOT) & 1u);
                const float* xs = xn_ring + slot * W_N + half * W_NH;
                ++tcount;
                Q1_LAP(0);
                const uint32_t buf = tile & (W_NBUF - 1);
                mbar_wait(bar_tfull + 8 * buf, (tile / W_NBUF) & 1u);
                tc_fence_after();
                Q1_LAP(2);
                const uint32_t taddr = tmem_base + W_ACC_COL + buf * W_N + half * W_NH + lane_taddr;
                const uint32_t pos0 = rt + half * W_NH;
                if (!(p.debug & 1u)) {
                    // this half of the query's accumulator row in two tensor-memory loads, one wait
                    uint32_t acc0[16], acc1[16];
                    if (p.debug & 4u) {   // timing experiment: no tensor-memory loads
#pragma unroll
                        for (int j = 0; j < 16; ++j) { acc0[j] = 0u; acc1[j] = 0u; }
                    } else {
                        tmem_ld16(taddr, acc0);
                        tmem_ld16(taddr + 16, acc1);
                        tmem_ld_wait();
                    }
                    Q1_LAP(1);
                    st_chunks += 2;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        float v[16];
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 t4 = *reinterpret_cast<const float4*>(xs + 16 * g + 4 * j4);
                            v[4 * j4 + 0] = t4.x; v[4 * j4 + 1] = t4.y; v[4 * j4 + 2] = t4.z; v[4 * j4 + 3] = t4.w;
                        }
                        // branch-free compare of 16 rows into a bit mask, then one loop over the set bits:
                        // every lane inserts into its own heap, the warp pays max-over-lanes(passing rows) rounds
                        uint32_t pass = 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (METRIC == METRIC_L2) v[j] = fmaf(-2.0f, __uint_as_float(g ? acc1[j] : acc0[j]), v[j]);   // |x|^2 - 2 q.x
                            else v[j] = __uint_as_float(g ? acc1[j] : acc0[j]) * v[j];                                    // -(q.x) [/ |x|]
                            pass |= (v[j] < thrp) ? (1u << j) : 0u;
                        }
                        Q1_LAP(3);
                        if (SMALL) {
                            // lane-divergent: every lane walks its own passing rows (rare after the first tiles)
                            while (pass) {
                                const uint32_t b = (uint32_t)__ffs((int)pass) - 1u;
                                pass &= pass - 1u;
                                float vb = v[0];
#pragma unroll
                                for (int j = 1; j < 16; ++j) vb = (b == (uint32_t)j) ? v[j] : vb;
                                if (vb < thrp) {
                                    ++st_app;
                                    unsigned long long k64 = ((unsigned long long)__float_as_uint(fmaxf(fmaf(vb, qa, qn), 0.0f)) << 32) |
                                                             (unsigned long long)(pos0 + 16u * (uint32_t)g + b);
#pragma unroll
                                    for (int i = 0; i < W_SMALL_N; ++i) {   // sorted insertion: the key sinks to its place
                                        const unsigned long long lo = best[i] < k64 ? best[i] : k64;
                                        k64 = best[i] < k64 ? k64 : best[i];
                                        best[i] = lo;
                                    }
                                    const uint32_t wb = (uint32_t)(best[W_SMALL_N - 1] >> 32);
                                    if (wb < F32_INF_BITS) {
                                        root = wb;
                                        thrp = fminf(thrp, (__uint_as_float(wb) - qn) * qia);
                                    }
                                }
                            }
                        } else
                        if (__any_sync(0xffffffffu, pass != 0)) {
                            ++st_rounds;
                            // filling phase: passing rows go straight into the list, in arrival order
                            if (__any_sync(0xffffffffu, pass != 0 && hcnt < (uint32_t)TC_KP)) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    if (((pass >> j) & 1u) && hcnt < (uint32_t)TC_KP) {
                                        ++hcnt;
                                        ++st_app;
                                        sts_v2(hb + 8 * hcnt, pos0 + 16u * (uint32_t)g + (uint32_t)j,
                                               __float_as_uint(fmaxf(fmaf(v[j], qa, qn), 0.0f)));
                                        pass &= ~(1u << j);
                                    }
                                }
                                const bool filled = hcnt == (uint32_t)TC_KP && root == 0xFFFFFFFFu;
                                Q1_LAP(4);
                                if (__any_sync(0xffffffffu, filled)) {
                                    const uint32_t r = wide_heapify(hb, filled);
                                    if (filled) {
                                        root = r;
                                        thrp = fminf(thrp, (__uint_as_float(root) - qn) * qia);
                                    }
                                    Q1_LAP(5);
                                }
                            }
                            // heap phase, warp-uniform loop: a lane parks one passing row per round; as soon as
                            // ANY lane's park is full every lane folds what it holds (lane by lane, each when its
                            // own park fills, would serialise the warp: one lane folds, 31 wait, every round)
                            while (__any_sync(0xffffffffu, pass != 0)) {
                                if (pass) {
                                    const uint32_t b = (uint32_t)__ffs((int)pass) - 1u;
                                    pass &= pass - 1u;
                                    float vb = v[0];
#pragma unroll
                                    for (int j = 1; j < 16; ++j) vb = (b == (uint32_t)j) ? v[j] : vb;
                                    if (vb < thrp) {         // (the threshold may have moved since the compare pass)
                                        ++st_app;
                                        sts_v2(pb + 8 * pcnt, pos0 + 16u * (uint32_t)g + b, __float_as_uint(fmaxf(fmaf(vb, qa, qn), 0.0f)));
                                        ++pcnt;
                                    }
                                }
                                if (__any_sync(0xffffffffu, pcnt == (uint32_t)W_PEND)) { Q1_LAP(6); fold(); Q1_LAP(7); }
                            }
                        }
                        Q1_LAP(6);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {   // accumulator may be overwritten: released on the LEADER's barrier
                    if (leader) mbar_arrive(bar_tempty + 8 * buf);
                    else mbar_arrive_remote(l_tempty + 8 * buf);
                    mbar_arrive(bar_nempty + 8 * slot);
                }
                Q1_LAP(0);
                if ((tile % W_FLUSH_TILES) == W_FLUSH_TILES - 1 && __any_sync(0xffffffffu, pcnt != 0)) fold();
                Q1_LAP(7);
                // any 32 rows below a value bound the query's global 32nd: share an improved root, and
                // every other tile pick up what the others published
                if (have && p.thr_g) {
                    if (root < root_pub) { root_pub = root; publish_bound(p, qi, root, peer_sent); }
                    thr_pending = *(volatile uint32_t*)(p.thr_g + qi);
                }
                ++tile;
            }
            // ---- item epilogue: publish the heaps as they are (unsorted rows; the shortlist merge sorts) ----
            Q1_LAP(0);
            if (__any_sync(0xffffffffu, pcnt != 0)) fold();
            Q1_LAP(7);
            if (have && p.thr_g && root < root_pub) publish_bound(p, qi, root, peer_sent);
            if (SMALL) {
                hcnt = 0;
#pragma unroll
                for (int i = 0; i < W_SMALL_N; ++i)
                    if (best[i] != KEY_NONE) {
                        sts_v2(hb + 8u * (uint32_t)(i + 1), (uint32_t)best[i], (uint32_t)(best[i] >> 32));
                        hcnt = (uint32_t)(i + 1);
                    }
            }
            __syncwarp();
            unsigned todo = __ballot_sync(0xffffffffu, have && hcnt > 0);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const uint32_t n = __shfl_sync(0xffffffffu, hcnt, src);
                const uint32_t qis = __shfl_sync(0xffffffffu, qi, src);
                const uint32_t sls = __shfl_sync(0xffffffffu, sl, src);
                const uint32_t hm = half * W_M + quarter * 32u + (uint32_t)src;
                uint64_t key = KEY_NONE;   // entry {row, d2 bits} read as one u64 = (d2 bits << 32) | row
                if ((uint32_t)lane < n) key = lds_u64(smem_u32(heaps) + hm * W_HEAP_BYTES + 8u * ((uint32_t)lane + 1u));
                // p.S = shortlist rows per (query, probe) = 2 * row ranges: row (2 * range + half)
                const size_t prow = ((size_t)qis * p.P + sls) * p.S + (size_t)W_SUBROWS * it.sub + half;
                p.partial[prow * TC_KP + lane] = key;
                if (lane == 0) p.row_stamp[prow] = p.stamp;
            }
            __syncwarp();   // the heaps may be refilled by the next item
            Q1_LAP(0);
        }
        if (warp == 4) Q1_LAP_DUMP(2);
        if (p.prof && warp == 4) {
            const uint32_t a = __reduce_add_sync(0xffffffffu, st_app);
            if (lane == 0) {
                unsigned long long* d = p.prof + ((size_t)blockIdx.x * 6 + 4) * 8;
                d[0] = a; d[1] = 0; d[2] = st_rounds; d[3] = 0; d[4] = tile; d[5] = st_chunks;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // no CTA may free tensor memory / exit while its peer's MMAs or arrivals are in flight
    if (p.prof && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.prof[((size_t)blockIdx.x * 6 + 3) * 8 + 5] = gt;
        p.prof[((size_t)blockIdx.x * 6 + 3) * 8 + 6] = (unsigned long long)clock64() - p.prof[((size_t)blockIdx.x * 6 + 3) * 8 + 6];
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)W_TMEM_COLS)
                     : "memory");
    }
}
