// tc_scan_pair.cuh — kernel P: the posting-list scan on CTA PAIRS (tcgen05 cta_group::2).
// Included by tc_scan.cu inside its anonymous namespace (shares TcScanParams, ScanItem, the warp
// sorting primitives and the stopwatch macros with kernel R).
//
// Why: kernel R holds the whole query tile of an item in shared memory (64 queries x D x 4 B =
// 96 KB at D = 384), so a list probed by more than 64 queries is a second item and its rows are
// streamed through the SM again (from L2).  The cost of an MMA does not depend on N and the
// epilogue's compare pass is free, so a 128-query item costs what a 64-query item
// costs — if its query tile fits.  With a two-SM MMA it does: the B operand (queries) is split
// between the two CTAs of a cluster, 64 queries each; each CTA streams its OWN 128-row tiles (A
// operand, M = 256 over the pair) and receives [its 128 rows x all N queries] in its own tensor
// memory.  Every list probed by <= 128 queries is read once.
//
// Roles per CTA as in kernel R (warp 0 producer, warp 1 MMA, warps 2-5 epilogue, warps 6-9 query
// loaders); differences:
//   * the LEADER (cluster rank 0) claims the work items and writes them into both CTAs' item rings;
//   * both producers issue TMA for their own tile of a 256-row super-tile; every load completes on
//     the leader's `full` barrier (the leader's producer arms it with the bytes of both);
//   * only the leader's MMA warp issues (cta_group::2, M = 256); its commits are multicast to the
//     `empty` / `tfull` / `qfree` barriers of both CTAs;
//   * both CTAs' epilogue warps release an accumulator on the leader's `tempty` barrier; both CTAs'
//     loaders announce their half of the query tile on the leader's `qready` barrier.
#pragma once

#ifndef FVDB_P2_NQ
#define FVDB_P2_NQ 96
#endif
constexpr int P2_NQ = FVDB_P2_NQ;                // queries per item (MMA N max): 96 leaves room for a 6-stage ring
constexpr int P2_NQH = P2_NQ / 2;                // ... of which each CTA stages half
constexpr int P2_QBLK_BYTES = P2_NQH * 128;      // bytes per k-block of this CTA's half of the query tile (6 KB)
constexpr int P2_CAP = (P2_NQ > 96) ? 16 : 32;   // pending candidates per query
constexpr int P2_FLUSH = P2_CAP / 2;             // merge a query once it holds more pending than this
constexpr int P2_NBUF = 4;                       // accumulator buffers ...
constexpr int P2_TSTRIDE = 128;                  // ... 128 tensor-memory columns apart
static_assert(P2_NQ % 32 == 0 && P2_NQ <= 128 && P2_QBLK_BYTES % 1024 == 0, "pair tile");

// ---- PTX of the pair ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// address of the same shared-memory location in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_cl(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait with cluster-scope acquire (arrivals may come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cl(uint32_t bar, uint32_t parity) {
    if (mbar_try_cl(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_cl(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// TMA tile load into THIS CTA's shared memory, completing on a barrier of the leader CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tmap, uint32_t bar_cluster, int c0,
                                                 int c1, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(tmap), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}

struct P2Smem {
    unsigned char* q_tile;    // KB x 8 KB: this CTA's half of the query tile
    unsigned char* ring;      // STAGES x 16 KB
    uint64_t* sorted;         // [P2_NQ][TC_KP]
    uint64_t* pend;           // [P2_NQ][P2_CAP]
    float* xn_ring;           // [R2_NSLOT][R2_ROWS]
    float* thrp;              // [P2_NQ]
    uint32_t* pcnt;           // [P2_NQ]
    uint32_t* qidx;           // [2][P2_NQ]
    uint32_t* qslot;          // [2][P2_NQ]
    float* qn;                // [2][P2_NQ]
    uint32_t* redo;           // [4]
    uint64_t* bars;
    uint32_t* tmem_ptr;
    uint32_t* sched;          // [TC_SCHED]
};
constexpr int P2_NBARS_FIXED = 2 * P2_NBUF + R2_NSLOT + 2 * TC_SCHED + 4;   // + 2 * STAGES

__device__ __forceinline__ P2Smem p2_carve(unsigned char* smem, uint32_t KB, uint32_t STAGES) {
    P2Smem m;
    m.q_tile = smem;
    m.ring = m.q_tile + (size_t)KB * P2_QBLK_BYTES;
    m.sorted = reinterpret_cast<uint64_t*>(m.ring + (size_t)STAGES * R2_STAGE_BYTES);
    m.pend = m.sorted + P2_NQ * TC_KP;
    m.xn_ring = reinterpret_cast<float*>(m.pend + P2_NQ * P2_CAP);
    m.thrp = m.xn_ring + R2_NSLOT * R2_ROWS;
    m.pcnt = reinterpret_cast<uint32_t*>(m.thrp + P2_NQ);
    m.qidx = m.pcnt + P2_NQ;
    m.qslot = m.qidx + 2 * P2_NQ;
    m.qn = reinterpret_cast<float*>(m.qslot + 2 * P2_NQ);
    m.redo = reinterpret_cast<uint32_t*>(m.qn + 2 * P2_NQ);
    m.bars = reinterpret_cast<uint64_t*>(m.redo + 4);
    m.tmem_ptr = reinterpret_cast<uint32_t*>(m.bars + 2 * STAGES + P2_NBARS_FIXED);
    m.sched = m.tmem_ptr + 1;
    return m;
}

size_t tc_scan_pair_smem_bytes(uint32_t KB, uint32_t stages) {
    return (size_t)KB * P2_QBLK_BYTES + (size_t)stages * R2_STAGE_BYTES + (size_t)P2_NQ * TC_KP * 8 +
           (size_t)P2_NQ * P2_CAP * 8 + (size_t)R2_NSLOT * R2_ROWS * 4 + (size_t)P2_NQ * 8 * 4 + 16 +
           (size_t)(2 * stages + P2_NBARS_FIXED) * 8 + 16 + (size_t)TC_SCHED * 4;
}

// Merge the pending candidates of up to four queries (owned by this warp) into their sorted
// shortlists, four 32-lane networks advanced in lockstep (the shuffle chains of one network are
// latency bound); lane i holds entry i.  n[g] pending entries (already clamped to P2_CAP).
__device__ __forceinline__ void p2_merge4(const P2Smem& sm, const uint32_t (&qs)[4], const uint32_t (&n)[4],
                                          const bool (&act)[4], uint64_t (&lst)[4], int lane) {
    uint64_t nw[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        lst[g] = KEY_NONE;
        nw[g] = KEY_NONE;
        if (act[g]) {
            lst[g] = sm.sorted[qs[g] * TC_KP + lane];
            if ((uint32_t)lane < n[g]) nw[g] = sm.pend[qs[g] * P2_CAP + lane];
        }
    }
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool keep_min = (((lane & k) == 0) == ((lane & j) == 0));
            uint64_t o[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) o[g] = shfl_xor64(nw[g], j);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint64_t mn = nw[g] < o[g] ? nw[g] : o[g], mx = nw[g] < o[g] ? o[g] : nw[g];
                nw[g] = keep_min ? mn : mx;
            }
        }
    }
    warp_merge32x4(lst, nw, lane);
}

__global__ void __launch_bounds__(R2_THREADS, 1)
tc_scan_pair_kernel(const __grid_constant__ CUtensorMap tmap, const TcScanParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t KB = p.KB;
    const uint32_t STAGES = p.stages;
    const P2Smem sm = p2_carve(smem, KB, STAGES);
    const uint32_t rank = cluster_ctarank();     // 0 = leader
    const bool leader = rank == 0;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = smem_u32(sm.bars);               // used in the leader only
    const uint32_t bar_empty = bar_full + 8 * STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * STAGES;
    const uint32_t bar_tempty = bar_tfull + 8 * P2_NBUF;        // leader only
    const uint32_t bar_nfull = bar_tempty + 8 * P2_NBUF;
    const uint32_t bar_sfull = bar_nfull + 8 * R2_NSLOT;
    const uint32_t bar_sempty = bar_sfull + 8 * TC_SCHED;       // leader only
    const uint32_t bar_qready = bar_sempty + 8 * TC_SCHED;      // leader only: both halves of the query tile staged
    const uint32_t bar_qfree = bar_qready + 8;                  // every MMA of the item has retired
    const uint32_t bar_mfree = bar_qfree + 8;                   // epilogue is done with an item's metadata
    const uint32_t bar_meta = bar_mfree + 8;                    // this CTA's copy of the item metadata staged
    // the leader's barriers as seen from this CTA
    const uint32_t l_full = mapa_u32(bar_full, 0), l_tempty = mapa_u32(bar_tempty, 0);
    const uint32_t l_sempty = mapa_u32(bar_sempty, 0), l_qready = mapa_u32(bar_qready, 0);

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int i = 0; i < P2_NBUF; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 8);    // four epilogue warps of each CTA
        }
        for (int i = 0; i < R2_NSLOT; ++i) mbar_init(bar_nfull + 8 * i, 1);
        for (int i = 0; i < TC_SCHED; ++i) {
            mbar_init(bar_sfull + 8 * i, 1);
            // leader: MMA lane + 4 epilogue + 4 loader warps; peer: producer + 4 epilogue + 4 loader warps
            mbar_init(bar_sempty + 8 * i, 18);
        }
        mbar_init(bar_qready, 256);
        mbar_init(bar_qfree, 1);
        mbar_init(bar_mfree, 4);
        mbar_init(bar_meta, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_ptr)),
                     "r"((uint32_t)R2_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // barriers of both CTAs initialised, tensor memory of both allocated
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_ptr;
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
        else mbar_arrive_cluster(l_sempty + 8 * ss);
    };

    if (warp == 0) {
        // ============ TMA producer (both CTAs) + tile scheduler (leader) ============
        uint32_t stage = 0, phase = 0, ss = 0, sphase = 0, tcount = 0;
        const uint64_t hint_first = 0x12F0000000000000ull;   // L2 evict-first: rows read once
        const uint64_t hint_normal = 0x1000000000000000ull;  // list shared by several items
        const uint32_t ring_base = smem_u32(sm.ring);
        const uint32_t peer_sched = mapa_u32(smem_u32(sm.sched), 1), peer_sfull = mapa_u32(bar_sfull, 1);
        while (true) {
            uint32_t item = 0;
            if (leader) {
                Q1_LAP(3);
                mbar_wait_cl(bar_sempty + 8 * ss, sphase ^ 1);
                if (lane == 0) {
                    item = atomicAdd(p.work_counter, 1u);
                    if (item >= n_items) item = ITEM_END;
                    sm.sched[ss] = item;
                    st_cluster_u32(peer_sched + 4 * ss, item);
                    mbar_arrive(bar_sfull + 8 * ss);
                    mbar_arrive_cluster(peer_sfull + 8 * ss);
                }
                item = __shfl_sync(0xffffffffu, item, 0);
            } else {
                mbar_wait_cl(bar_sfull + 8 * ss, sphase);
                item = *reinterpret_cast<volatile uint32_t*>(sm.sched + ss);
                __syncwarp();
                if (lane == 0) release_slot(ss);
            }
            if (++ss == TC_SCHED) { ss = 0; sphase ^= 1; }
            Q1_LAP(0);
            if (item == ITEM_END) break;
            const ScanItem it = p.items[item];
            if (it.pair_count == 0 || it.row_begin >= it.row_end) continue;
            lap[6] += 1;
            Q1_LAP(1);
            const uint64_t hint = (it.identity == 2 || (!it.identity && it.slot > 1)) ? hint_normal : hint_first;
            // super-tiles of 256 rows: this CTA's tile is rows [rt + 128 * rank, + 128)
            for (uint32_t rt = it.row_begin; rt < it.row_end; rt += 2 * R2_ROWS) {
                const uint32_t my_rt = rt + rank * R2_ROWS;
                const bool mine_live = my_rt < it.row_end;
                const bool peer_live = rt + R2_ROWS < it.row_end;     // the peer's (rank 1) tile holds rows of this list
                float xnv[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const uint32_t pos = my_rt + h * 32 + lane;
                    float xn = __uint_as_float(F32_INF_BITS);
                    if (pos < it.row_end) {
                        bool live = true;
                        if (p.tomb || p.filt) {
                            const uint32_t id = p.ids[pos];
                            if (p.tomb && bit_test(p.tomb, p.tomb_bits, id)) live = false;
                            else if (p.filt && !bit_test(p.filt, p.filt_bits, id)) live = false;
                        }
                        if (live) xn = __ldg(p.xnorm + pos);
                    }
                    xnv[h] = xn;
                }
                __syncwarp();
                lap[7] += 1;
                for (uint32_t kb = 0; kb < KB; ++kb) {
                    Q1_LAP(3);
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    Q1_LAP(2);
                    if (elect_one()) {
                        // the leader arms its barrier with the bytes of both CTAs' boxes; a tile past the
                        // end of the list is not loaded (its rows are masked by +inf norms)
                        if (leader)
                            mbar_expect_tx(bar_full + 8 * stage, (uint32_t)R2_STAGE_BYTES * (peer_live ? 2u : 1u));
                        if (mine_live)
                            tma_load_2d_pair(ring_base + stage * R2_STAGE_BYTES, &tmap, l_full + 8 * stage,
                                             (int)(kb * TC_KB_FLOATS), (int)my_rt, hint);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                const uint32_t slot = tcount % R2_NSLOT;
#pragma unroll
                for (int h = 0; h < 4; ++h) sm.xn_ring[slot * R2_ROWS + h * 32 + lane] = xnv[h];
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_nfull + 8 * slot);
                ++tcount;
            }
        }
        Q1_LAP_DUMP(0);
    } else if (warp == 1) {
        // ============ MMA issuer: the leader's warp only ============
        if (leader) {
            uint32_t stage = 0, phase = 0, ss = 0, sphase = 0, tile = 0, nit = 0;
            const uint32_t q_base = smem_u32(sm.q_tile);
            const uint32_t ring_base = smem_u32(sm.ring);
            while (true) {
                Q1_LAP(4);
                mbar_wait(bar_sfull + 8 * ss, sphase);
                const uint32_t item = sm.sched[ss];
                __syncwarp();
                if (lane == 0) release_slot(ss);
                if (++ss == TC_SCHED) { ss = 0; sphase ^= 1; }
                Q1_LAP(0);
                if (item == ITEM_END) break;
                const ScanItem it = p.items[item];
                if (it.pair_count == 0 || it.row_begin >= it.row_end) continue;
                const uint32_t ncols = (it.pair_count + 15u) & ~15u;
                const uint32_t idesc = umma_idesc_tf32(2 * R2_ROWS, ncols);   // M = 256 over the pair
                Q1_LAP(4);
                mbar_wait_cl(bar_qready, nit & 1u);
                tc_fence_after();
                Q1_LAP(1);
                for (uint32_t rt = it.row_begin; rt < it.row_end; rt += 2 * R2_ROWS) {
                    const uint32_t buf = tile & (P2_NBUF - 1);
                    Q1_LAP(4);
                    mbar_wait_cl(bar_tempty + 8 * buf, ((tile / P2_NBUF) & 1u) ^ 1u);
                    tc_fence_after();
                    Q1_LAP(2);
                    const uint32_t d_tmem = tmem_base + buf * P2_TSTRIDE;
                    const bool last_tile = rt + 2 * R2_ROWS >= it.row_end;
                    for (uint32_t kb = 0; kb < KB; ++kb) {
                        Q1_LAP(4);
                        mbar_wait_cl(bar_full + 8 * stage, phase);
                        tc_fence_after();
                        Q1_LAP(3);
                        if (elect_one()) {
                            const uint64_t a0 = umma_desc_sw128(ring_base + stage * R2_STAGE_BYTES);
                            const uint64_t b0 = umma_desc_sw128(q_base + kb * P2_QBLK_BYTES);
#pragma unroll
                            for (uint32_t k4 = 0; k4 < 4; ++k4)
                                if (!(p.debug & 8u))
                                    umma_tf32_pair(d_tmem, a0 + 2 * k4, b0 + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
                            umma_commit_pair(bar_empty + 8 * stage);
                            if (kb + 1 == KB) {
                                umma_commit_pair(bar_tfull + 8 * buf);
                                if (last_tile) umma_commit_pair(bar_qfree);
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
    } else if (warp >= 6) {
        // ============ query loaders (128 threads per CTA): metadata of the whole item, this CTA's half of the tile ============
        const int lw = warp - 6;
        const uint32_t D = p.D;
        uint32_t ss = 0, sphase = 0, nit = 0;
        while (true) {
            Q1_LAP(3);
            mbar_wait_cl(bar_sfull + 8 * ss, sphase);
            const uint32_t item = *reinterpret_cast<volatile uint32_t*>(sm.sched + ss);
            __syncwarp();
            if (lane == 0) release_slot(ss);
            if (++ss == TC_SCHED) { ss = 0; sphase ^= 1; }
            Q1_LAP(0);
            if (item == ITEM_END) break;
            const ScanItem it = p.items[item];
            if (it.pair_count == 0 || it.row_begin >= it.row_end) continue;
            const uint32_t cnt = it.pair_count;
            const uint32_t ncols = (cnt + 15u) & ~15u;
            const uint32_t half = ncols >> 1;            // queries staged by each CTA (multiple of 8)
            const uint32_t qbase = rank * half;          // first column of this CTA's half
            // lane l holds the query indices of local rows l and l + 32 of this CTA's half
            uint32_t qi0 = ID_NONE, qi1 = ID_NONE;
            {
                const uint32_t j0 = qbase + lane, j1 = qbase + lane + 32;
                if ((uint32_t)lane < half && j0 < cnt) qi0 = it.identity ? it.pair_begin + j0 : p.pair_q[it.pair_begin + j0];
                if ((uint32_t)lane + 32 < half && j1 < cnt) qi1 = it.identity ? it.pair_begin + j1 : p.pair_q[it.pair_begin + j1];
            }
            // metadata buffer (nit & 1) was last used by item nit - 2
            Q1_LAP(3);
            if (nit >= 2) mbar_wait(bar_mfree, nit & 1u);
            Q1_LAP(1);
            if (lw == 0) {
                const uint32_t mb = (nit & 1u) * P2_NQ;
#pragma unroll
                for (int g = 0; g < P2_NQ / 32; ++g) {
                    const uint32_t j = (uint32_t)lane + 32u * g;
                    uint32_t qi = ID_NONE, sl = 0;
                    if (j < cnt) {
                        if (it.identity) { qi = it.pair_begin + j; sl = it.slot; }
                        else { qi = p.pair_q[it.pair_begin + j]; sl = p.pair_slot[it.pair_begin + j]; }
                    }
                    sm.qidx[mb + j] = qi;
                    sm.qslot[mb + j] = sl;
                    sm.qn[mb + j] = (qi != ID_NONE) ? p.qnorm[qi] : 0.f;
                }
            }
            mbar_arrive(bar_meta);   // the epilogue of this CTA may start the item
            // the query tile may be overwritten once every MMA of the previous item has retired
            Q1_LAP(3);
            if (nit >= 1) mbar_wait(bar_qfree, (nit - 1) & 1u);
            Q1_LAP(2);
            {
                const uint32_t f4_per_row = KB * 8;
                for (uint32_t q0 = lw; q0 < half; q0 += 16) {
                    const float4* src[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t q = q0 + 4 * g;
                        const uint32_t qa = __shfl_sync(0xffffffffu, qi0, q & 31), qb = __shfl_sync(0xffffffffu, qi1, q & 31);
                        const uint32_t qi = (q < 32) ? qa : qb;
                        src[g] = (q < half && qi != ID_NONE) ? reinterpret_cast<const float4*>(p.Q + (size_t)qi * D) : nullptr;
                    }
                    for (uint32_t c = lane; c < f4_per_row; c += 96) {
                        float4 v[3][4];
#pragma unroll
                        for (int u = 0; u < 3; ++u)
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                v[u][g] = (src[g] && c + 32 * u < f4_per_row) ? __ldg(src[g] + c + 32 * u)
                                                                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            const uint32_t cc = c + 32 * u;
                            if (cc >= f4_per_row) break;
                            const uint32_t kb = cc >> 3, ch = cc & 7;
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const uint32_t q = q0 + 4 * g;
                                if (q < half)
                                    *reinterpret_cast<float4*>(sm.q_tile + (size_t)kb * P2_QBLK_BYTES + q * 128 +
                                                               ((ch ^ (q & 7)) << 4)) = v[u][g];
                            }
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA of the pair)
            if (leader) mbar_arrive(bar_qready);
            else mbar_arrive_cluster(l_qready);
            ++nit;
        }
        if (warp == 6 && p.prof && lane == 0) { for (int i_ = 0; i_ < 4; ++i_) p.prof[((size_t)blockIdx.x * 6 + 3) * 8 + i_] = lap[i_]; }
    } else {
        // ================= epilogue (warps 2-5): one thread = one row of this CTA's tile =================
        const int ew = warp - 2;                 // owner stripe: this warp merges queries j with j % 4 == ew
        const int quarter = warp & 3;            // TMEM lane quarter this warp may read
        const int trow = quarter * 32 + lane;    // row within the tile == TMEM lane
        const int et = ew * 32 + lane;           // 0..127
        const uint32_t lane_taddr = (uint32_t)(quarter * 32) << 16;
        uint32_t ss = 0, sphase = 0, tile = 0, nit = 0;
        uint32_t st_app = 0, st_ovf = 0, st_merge = 0, st_replay = 0, st_chunks = 0;
        // the query this lane owns in the merge phases: j = lane * 4 + ew  (32 lanes x 4 warps = 128)
        const uint32_t jown = (uint32_t)lane * 4u + (uint32_t)ew;
        while (true) {
            Q1_LAP(6);
            mbar_wait_cl(bar_sfull + 8 * ss, sphase);
            const uint32_t item = *reinterpret_cast<volatile uint32_t*>(sm.sched + ss);
            __syncwarp();
            if (lane == 0) release_slot(ss);
            if (++ss == TC_SCHED) { ss = 0; sphase ^= 1; }
            Q1_LAP(0);
            if (item == ITEM_END) break;
            const ScanItem it = p.items[item];
            if (it.pair_count == 0 || it.row_begin >= it.row_end) continue;
            const uint32_t cnt = it.pair_count;
            const uint32_t ncols = (cnt + 15u) & ~15u;
            const uint32_t mb = (nit & 1u) * P2_NQ;
            mbar_wait(bar_meta, nit & 1u);             // metadata of this item is staged
            Q1_LAP(1);
            if (nit >= 1) {                            // ... and only now release the previous item's
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_mfree);
            }
            // ---- item prologue: empty shortlists, thresholds from the shared bounds ----
            if (et < P2_NQ) {
                float thr = -__uint_as_float(F32_INF_BITS);  // padded columns never pass
                if ((uint32_t)et < cnt) {
                    const uint32_t g = p.thr_g ? *(volatile uint32_t*)(p.thr_g + sm.qidx[mb + et]) : (uint32_t)0x7f800000u;
                    thr = __uint_as_float(g) - sm.qn[mb + et];
                }
                sm.thrp[et] = thr;
                sm.pcnt[et] = 0;
            }
            if (et < 4) sm.redo[et] = 0;
            for (int i = et; i < P2_NQ * TC_KP; i += 128) sm.sorted[i] = KEY_NONE;
            const bool own = jown < cnt;
            const uint32_t qi_own = own ? sm.qidx[mb + jown] : 0u;
            const float qn_own = own ? sm.qn[mb + jown] : 0.f;
            uint32_t thr_pending = F32_INF_BITS;
            uint32_t peer_sent = F32_INF_BITS;
            epi_bar_n(1);
            Q1_LAP(6);

            // merge every owned query selected by `need` (warp-uniform mask over the 32 lanes)
            auto merge_owned = [&](unsigned need) {
                st_merge += __popc(need);
                if (__popc(need) == 1) {
                    // the common case: one query to fold -> a single 32-lane network
                    const int src = __ffs(need) - 1;
                    const uint32_t q = (uint32_t)src * 4u + (uint32_t)ew;
                    const uint32_t c = sm.pcnt[q];
                    const uint32_t n = min(c, (uint32_t)P2_CAP);
                    uint64_t nw = ((uint32_t)lane < n) ? sm.pend[q * P2_CAP + lane] : KEY_NONE;
                    uint64_t lst = sm.sorted[q * TC_KP + lane];
                    nw = warp_sort32(nw, lane);
                    lst = warp_merge32(lst, nw, lane);
                    sm.sorted[q * TC_KP + lane] = lst;
                    const uint64_t last = shfl64(lst, 31);
                    if (lane == src) {
                        sm.pcnt[q] = 0;
                        if (last != KEY_NONE) {
                            sm.thrp[q] = fminf(sm.thrp[q], __uint_as_float((uint32_t)(last >> 32)) - qn_own);
                            if (p.thr_g) publish_bound(p, qi_own, (uint32_t)(last >> 32), peer_sent);
                        }
                        if (c > (uint32_t)P2_CAP) atomicOr(&sm.redo[q >> 5], 1u << (q & 31));
                    }
                    __syncwarp();
                    return;
                }
                while (need) {
                    uint32_t qs[4], nn[4], over = 0;
                    bool act[4];
                    int src[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        act[g] = need != 0;
                        src[g] = act[g] ? (__ffs(need) - 1) : 0;
                        if (act[g]) need &= need - 1;
                        qs[g] = (uint32_t)src[g] * 4u + (uint32_t)ew;
                        const uint32_t c = act[g] ? sm.pcnt[qs[g]] : 0u;
                        nn[g] = min(c, (uint32_t)P2_CAP);
                        if (c > (uint32_t)P2_CAP) over |= 1u << g;
                    }
                    uint64_t lst[4];
                    p2_merge4(sm, qs, nn, act, lst, lane);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (!act[g]) continue;
                        sm.sorted[qs[g] * TC_KP + lane] = lst[g];
                        const uint64_t last = shfl64(lst[g], 31);
                        if (lane == src[g]) {
                            sm.pcnt[qs[g]] = 0;
                            if (last != KEY_NONE) {
                                sm.thrp[qs[g]] = fminf(sm.thrp[qs[g]], __uint_as_float((uint32_t)(last >> 32)) - qn_own);
                                if (p.thr_g) publish_bound(p, qi_own, (uint32_t)(last >> 32), peer_sent);
                            }
                            if (over & (1u << g)) atomicOr(&sm.redo[qs[g] >> 5], 1u << (qs[g] & 31));
                        }
                    }
                    __syncwarp();
                }
            };

            // ---- row tiles (this CTA's half of every 256-row super-tile) ----
            for (uint32_t rt = it.row_begin; rt < it.row_end; rt += 2 * R2_ROWS) {
                const uint32_t slot = tile % R2_NSLOT;
                Q1_LAP(4);
                mbar_wait(bar_nfull + 8 * slot, (tile / R2_NSLOT) & 1u);
                const float xn = sm.xn_ring[slot * R2_ROWS + trow];
                const uint32_t buf = tile & (P2_NBUF - 1);
                mbar_wait(bar_tfull + 8 * buf, (tile / P2_NBUF) & 1u);
                tc_fence_after();
                Q1_LAP(2);
                const uint32_t taddr = tmem_base + buf * P2_TSTRIDE + lane_taddr;
                const uint32_t pos = rt + rank * R2_ROWS + (uint32_t)trow;
                uint64_t ovf0 = 0, ovf1 = 0;  // queries (0..63 / 64..127) whose pending list was full when this row passed

                auto append = [&](uint32_t q, float v) {
                    const uint32_t s = atomicAdd(&sm.pcnt[q], 1u);
                    ++st_app;
                    if (s < (uint32_t)P2_CAP)
                        sm.pend[q * P2_CAP + s] =
                            ((uint64_t)__float_as_uint(fmaxf(v + sm.qn[mb + q], 0.0f)) << 32) | (uint64_t)pos;
                    else {
                        if (q < 64) ovf0 |= 1ull << q; else ovf1 |= 1ull << (q - 64);
                        ++st_ovf;
                    }
                };

                // 16 columns at a time; the tensor-memory load of the NEXT chunk is in flight while this
                // one is compared (one compact copy of the compare/append code: the instruction cache
                // matters more here than the unrolling)
                const uint32_t nch = (p.debug & 1u) ? 0u : (ncols >> 4);
                uint32_t cur[16], nxt[16];
                if (nch) { tmem_ld16(taddr, cur); tmem_ld_wait(); }
#pragma unroll 1
                for (uint32_t c = 0; c < nch; ++c) {
                    const uint32_t c0 = 16 * c;
                    ++st_chunks;
                    if (c + 1 < nch) tmem_ld16(taddr + c0 + 16, nxt);
                    float thr[16];
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 t4 = *reinterpret_cast<const float4*>(sm.thrp + c0 + 4 * j4);
                        thr[4 * j4 + 0] = t4.x; thr[4 * j4 + 1] = t4.y; thr[4 * j4 + 2] = t4.z; thr[4 * j4 + 3] = t4.w;
                    }
                    uint32_t pass = 0;   // branch-free compare, then one loop over the set bits (see kernel R)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float v = fmaf(-2.0f, __uint_as_float(cur[j]), xn);  // |x|^2 - 2 x.q
                        pass |= (v < thr[j]) ? (1u << j) : 0u;
                    }
                    while (pass) {
                        const uint32_t b = (uint32_t)__ffs((int)pass) - 1u;
                        pass &= pass - 1u;
                        uint32_t a = cur[0];
#pragma unroll
                        for (int j = 1; j < 16; ++j) a = (b == (uint32_t)j) ? cur[j] : a;
                        append(c0 + b, fmaf(-2.0f, __uint_as_float(a), xn));
                    }
                    if (c + 1 < nch) {
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) cur[j] = nxt[j];
                    }
                }
                Q1_LAP(3);
                epi_bar_n(2);  // every candidate of this tile is in the pending lists
                // ---- merge phase: owners fold long pending lists, refresh shared thresholds ----
                while (true) {
                    if (own) {
                        sm.thrp[jown] = fminf(sm.thrp[jown], __uint_as_float(thr_pending) - qn_own);
                        if (p.thr_g) thr_pending = *(volatile uint32_t*)(p.thr_g + qi_own);
                    }
                    const uint32_t pc = own ? sm.pcnt[jown] : 0u;
                    merge_owned(__ballot_sync(0xffffffffu, pc > (uint32_t)P2_FLUSH));
                    epi_bar_n(1);  // thresholds / pending counters settled
                    const uint32_t r0 = sm.redo[0], r1 = sm.redo[1], r2 = sm.redo[2], r3 = sm.redo[3];
                    if ((r0 | r1 | r2 | r3) == 0) break;
                    ++st_replay;
                    // replay: rows that met a full pending list are tested against the new thresholds
                    epi_bar_n(2);
                    if (et < 4) sm.redo[et] = 0;
#pragma unroll
                    for (int w2 = 0; w2 < 2; ++w2) {
                        uint64_t rm = w2 ? (((uint64_t)r3 << 32) | r2) : (((uint64_t)r1 << 32) | r0);
                        uint64_t& ovf = w2 ? ovf1 : ovf0;
                        while (rm) {
                            const uint32_t b = (uint32_t)__ffsll((long long)rm) - 1u;
                            rm &= rm - 1;
                            const uint32_t q = b + 64u * w2;
                            uint32_t a;
                            tmem_ld1(taddr + q, a);
                            tmem_ld_wait();
                            if ((ovf >> b) & 1ull) {
                                ovf &= ~(1ull << b);
                                const float v = fmaf(-2.0f, __uint_as_float(a), xn);
                                if (v < sm.thrp[q]) append(q, v);
                            }
                        }
                    }
                    epi_bar_n(1);
                    // every replayed query is merged again (its pending list may have refilled)
                    {
                        const uint32_t rw = (jown < 32) ? r0 : (jown < 64) ? r1 : (jown < 96) ? r2 : r3;
                        const bool mine = own && ((rw >> (jown & 31)) & 1u);
                        merge_owned(__ballot_sync(0xffffffffu, mine && sm.pcnt[jown] > 0));
                    }
                    epi_bar_n(2);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {   // accumulator may be overwritten: released on the LEADER's barrier
                    if (leader) mbar_arrive(bar_tempty + 8 * buf);
                    else mbar_arrive_cluster(l_tempty + 8 * buf);
                }
                ++tile;
            }
            // ---- item epilogue: fold what is pending, publish the shortlists ----
            Q1_LAP(4);
            {
                const uint32_t pc = own ? sm.pcnt[jown] : 0u;
                const bool has_sorted = own && sm.sorted[jown * TC_KP] != KEY_NONE;
                merge_owned(__ballot_sync(0xffffffffu, pc > 0 && has_sorted));
                __syncwarp();
                // both CTAs of the pair hold a shortlist of the same (query, probe) over their own rows:
                // `partial` has 2 P rows per query, the leader publishes into row [slot], the peer into
                // row [P + slot]; the shortlist merge after the scan folds all of them
                for (uint32_t j = (uint32_t)ew; j < cnt; j += 4) {
                    uint64_t mine = sm.sorted[j * TC_KP + lane];
                    if (__shfl_sync(0xffffffffu, mine == KEY_NONE ? 1 : 0, 0)) {   // no sorted list: pending, as is
                        const uint32_t n = sm.pcnt[j];
                        if (n == 0) continue;                                      // partial is pre-filled
                        mine = ((uint32_t)lane < n) ? sm.pend[j * P2_CAP + lane] : KEY_NONE;
                    }
                    const size_t prow = ((size_t)sm.qidx[mb + j] * (2u * p.P) + rank * p.P + sm.qslot[mb + j]) * (p.S ? p.S : 1u) + it.sub;
                    p.partial[prow * TC_KP + lane] = mine;
                    if (lane == 0) p.row_stamp[prow] = p.stamp;
                }
            }
            epi_bar_n(1);  // pools may be re-initialised for the next item
            ++nit;
            Q1_LAP(5);
        }
        if (warp == 2) Q1_LAP_DUMP(2);
        if (p.prof && warp == 2) {
            const uint32_t a = __reduce_add_sync(0xffffffffu, st_app), o = __reduce_add_sync(0xffffffffu, st_ovf);
            if (lane == 0) {
                unsigned long long* d = p.prof + ((size_t)blockIdx.x * 6 + 4) * 8;
                d[0] = a; d[1] = o; d[2] = st_merge; d[3] = st_replay; d[4] = tile; d[5] = st_chunks;
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
                     "r"((uint32_t)R2_TMEM_COLS)
                     : "memory");
    }
}
