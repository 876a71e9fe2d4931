// seglog.cuh — HBM-resident segmented entry buffer with asynchronous pinned-host spill.
//
// SURVEY.md §8(f)-1 / north_star: on the hot path the RocksDB-backed RaftLog
// (M/command/storage/RocksLog.java:82-242, one RocksDB per context, key = 8-byte BE index,
// value = 8-byte BE term || Kryo(cmd), flushWal(true) on every write) is replaced by an in-memory
// store: entry payloads live in a ring of fixed-size SEGMENTS in HBM; a segment that must be recycled
// is first copied to pinned host memory on a side stream (the cold tier).  The index->term map the
// protocol needs stays in the run-length tables of the step kernel; this file only moves payload bytes.
//
//   RocksLog.newEntry / append  (put)          -> rafting_log_append  (batch: one H2D + one index kernel)
//   RocksLog.get / batch        (multiGet)     -> rafting_log_read    (one group), rafting_log_gather
//                                                 (many (gid, first, count) ranges = the AE plans of a step)
//   RocksLog.truncate / flush   (deleteRange)  -> visibility of a record is decided by the group's stored key range
//                                                 [g_lo, g_hi] in the tables, and a re-appended index overwrites
//                                                 its ring slot; rafting_log_trim reclaims what a flush left behind
//
// Record layout inside a segment: 24-byte header {gid u32, len u32, index i64, term i64} + payload padded
// to 8 bytes.  Records never straddle segments.  A logical byte offset (segment number * seg_bytes + offset)
// names a record forever; the HBM arena holds the newest `nseg` segments, older ones are read from the
// host copy.  Per group a device ring of K slots maps index -> logical offset for the gather kernel.
//
// Included at the end of engine.cu (same translation unit: it uses rafting_engine, fail(), CU()).
#pragma once

namespace rafting {

struct alignas(16) SegHdr { uint32_t gid, len; int64_t index, term; uint64_t seq; };
static_assert(sizeof(SegHdr) == 32, "segment record header");

struct HostLoc { uint64_t off; uint32_t len; int64_t term; };      // len == 0xffffffff: absent
struct GroupIdx { int64_t base = 0; std::vector<HostLoc> v; };     // host index of one group: v[i] <-> index base + i

struct SegLog {
    uint32_t seg_bytes = 0, nseg = 0, K = 0;
    uint8_t* arena = nullptr;                 // [nseg * seg_bytes] in HBM
    uint64_t head = 0;                        // logical offset of the next record
    uint64_t spilled_upto = 0;                // logical segments < this have a host copy enqueued
    unsigned long long* ring = nullptr;       // [G * K]: 1 + logical offset of the newest record whose index maps to the slot
    std::vector<uint8_t*> cold;               // pinned host copy per logical segment (null until spilled)
    std::vector<cudaEvent_t> cold_ready;      // spill completion per logical segment
    std::vector<GroupIdx> index;              // authoritative (lengths, terms, cold tier); the device ring is a cache of it
    uint8_t* stage = nullptr; size_t stage_cap = 0;       // pinned staging of one append batch
    void* d_req = nullptr; size_t d_req_cap = 0;           // device scratch (append records / gather requests)
    uint8_t* d_blob = nullptr; size_t d_blob_cap = 0;      // staged payload blob of the append in flight
    uint8_t* d_out = nullptr; size_t d_out_cap = 0;
    cudaStream_t s_spill = nullptr;
    uint64_t appended = 0, spilled_bytes = 0, hbm_hits = 0, cold_hits = 0, indexed = 0;
    std::vector<uint32_t> seg_live;           // live (indexed, not overwritten, not trimmed) records per logical segment
    uint64_t trimmed = 0, cold_freed_bytes = 0, spills_skipped = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;    // device time of the last gather's kernel
    float last_gather_kernel_ms = 0; uint64_t last_gather_bytes = 0;
};

struct AppendRec { uint32_t gid, len; int64_t index, term; uint64_t loc; uint64_t src; };   // 40 B per appended entry
// 8 lanes per appended record: write the header, copy the payload from the staged blob into the arena and
// publish the record in the group's ring.  Logical offsets only grow, so atomicMax makes "the latest put wins"
// (RocksDB semantics) hold inside a batch and across batches without any ordering between threads.
__global__ void seglog_scatter_kernel(const AppendRec* __restrict__ recs, uint32_t n, uint64_t seq0, const uint8_t* __restrict__ blob,
                                      uint8_t* __restrict__ arena, uint64_t arena_bytes, unsigned long long* ring, uint32_t K) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t en = t >> 3, sub = t & 7;
    if (en >= n) return;
    const AppendRec r = recs[en];
    uint8_t* dst = arena + (r.loc % arena_bytes);
    if (sub == 0) {
        SegHdr h; h.gid = r.gid; h.len = r.len; h.index = r.index; h.term = r.term; h.seq = seq0 + en;
        *(int4*)dst = *(const int4*)&h; *(int4*)(dst + 16) = *((const int4*)&h + 1);
        atomicMax(&ring[(size_t)r.gid * K + ((uint64_t)r.index & (K - 1))], (unsigned long long)(r.loc + 1));
    }
    const uint8_t* src = blob + r.src;
    uint8_t* pay = dst + sizeof(SegHdr);
    if ((r.src & 15) == 0) {
        const uint32_t words = r.len >> 4;
        for (uint32_t i = sub; i < words; i += 8) ((int4*)pay)[i] = ((const int4*)src)[i];
        for (uint32_t i = (words << 4) + sub; i < r.len; i += 8) pay[i] = src[i];
    } else {
        for (uint32_t i = sub; i < r.len; i += 8) pay[i] = src[i];
    }
}

struct GatherReq { uint32_t gid; uint32_t slot; int64_t index; uint64_t out_off; };   // slot: position in the reply list

// 8 lanes per requested entry (4 entries per warp, so four dependent chains ring -> header -> payload are in
// flight per warp), 16-byte vector copies HBM arena -> contiguous send buffer.  lens[e] = 0xffffffff when the
// ring does not point at a resident record of exactly (gid, index); the host then serves it from its index.
__global__ void seglog_gather_kernel(const GatherReq* __restrict__ req, uint32_t n, const unsigned long long* __restrict__ ring,
                                     uint32_t K, const uint8_t* __restrict__ arena, uint64_t arena_bytes,
                                     uint64_t oldest_resident, uint32_t* __restrict__ lens, uint8_t* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t en = t >> 3, sub = t & 7;
    if (en >= n) return;
    const GatherReq q = req[en];
    const unsigned long long loc1 = ring[(size_t)q.gid * K + ((uint64_t)q.index & (K - 1))];
    bool hit = loc1 != 0 && loc1 - 1 >= oldest_resident;
    uint32_t len = 0;
    const uint8_t* rec = arena + ((loc1 - 1) % arena_bytes);
    if (hit) {
        const int4 h0 = *(const int4*)rec;                       // gid, len, index
        const long long idx = ((long long)(unsigned)h0.w << 32) | (unsigned)h0.z;
        hit = (uint32_t)h0.x == q.gid && idx == q.index;
        len = (uint32_t)h0.y;
    }
    if (sub == 0) lens[en] = hit ? len : 0xffffffffu;
    if (!hit) return;
    const int4* src = (const int4*)(rec + sizeof(SegHdr));       // records and payloads are 16-byte aligned
    int4* dst = (int4*)(out + q.out_off);                        // out_off is 16-byte aligned
    const uint32_t words = (len + 15) >> 4;
    for (uint32_t i = sub; i < words; i += 8) dst[i] = src[i];
}

}  // namespace rafting

using rafting::SegLog; using rafting::SegHdr; using rafting::HostLoc; using rafting::GatherReq; using rafting::AppendRec;

static void seglog_release(rafting_engine* e) {
    SegLog* L = e->seglog; if (!L) return;
    if (L->s_spill) { cudaStreamSynchronize(L->s_spill); cudaStreamDestroy(L->s_spill); }
    for (auto ev : L->cold_ready) if (ev) cudaEventDestroy(ev);
    for (auto p : L->cold) if (p) cudaFreeHost(p);
    if (L->arena) cudaFree(L->arena);
    if (L->ring) cudaFree(L->ring);
    if (L->stage) cudaFreeHost(L->stage);
    if (L->d_req) cudaFree(L->d_req);
    if (L->d_out) cudaFree(L->d_out);
    if (L->d_blob) cudaFree(L->d_blob);
    if (L->t0) { cudaEventDestroy(L->t0); cudaEventDestroy(L->t1); }
    delete L; e->seglog = nullptr;
}

extern "C" int rafting_log_config(rafting_engine_t* e, uint32_t segment_bytes, uint32_t hbm_segments, uint32_t ring_slots) {
    if (!e) return fail(RAFTING_E_INVAL, "null argument");
    if (e->seglog) return fail(RAFTING_E_INVAL, "entry buffer already configured");
    if (segment_bytes < 4096 || (segment_bytes & 15) || hbm_segments < 2 || ring_slots == 0 || (ring_slots & (ring_slots - 1)))
        return fail(RAFTING_E_INVAL, "segment_bytes >= 4096 and 16-aligned, hbm_segments >= 2, ring_slots a power of two");
    CU(cudaSetDevice(e->cfg.device));
    SegLog* L = new SegLog();
    L->seg_bytes = segment_bytes; L->nseg = hbm_segments; L->K = ring_slots;
    e->seglog = L;
    L->index.resize(e->G);
    const size_t ring = (size_t)e->G * ring_slots;
    CU(cudaMalloc(&L->arena, (size_t)segment_bytes * hbm_segments));
    CU(cudaMalloc(&L->ring, ring * 8));
    CU(cudaMemset(L->ring, 0, ring * 8));                          // 0 = empty slot
    CU(cudaStreamCreateWithFlags(&L->s_spill, cudaStreamNonBlocking));
    if (!e->ev_seg) CU(cudaEventCreateWithFlags(&e->ev_seg, cudaEventDisableTiming));
    return RAFTING_OK;
}

static inline const HostLoc* seglog_find(const SegLog* L, uint32_t gid, int64_t index) {
    const rafting::GroupIdx& gi = L->index[gid];
    if (index < gi.base || index >= gi.base + (int64_t)gi.v.size()) return nullptr;
    const HostLoc& h = gi.v[(size_t)(index - gi.base)];
    return h.len == 0xffffffffu ? nullptr : &h;
}
static inline void seglog_put(SegLog* L, uint32_t gid, int64_t index, const HostLoc& hl) {
    rafting::GroupIdx& gi = L->index[gid];
    const HostLoc none = {0, 0xffffffffu, 0};
    if (gi.v.empty()) gi.base = index;
    if (index < gi.base) { gi.v.insert(gi.v.begin(), (size_t)(gi.base - index), none); gi.base = index; }
    if (index >= gi.base + (int64_t)gi.v.size()) gi.v.resize((size_t)(index - gi.base) + 1, none);
    HostLoc& slot = gi.v[(size_t)(index - gi.base)];
    if (slot.len == 0xffffffffu) L->indexed++;
    else L->seg_live[slot.off / L->seg_bytes]--;                   // the overwritten record is dead
    const uint64_t seg = hl.off / L->seg_bytes;
    if (L->seg_live.size() <= seg) L->seg_live.resize(seg + 1, 0);
    L->seg_live[seg]++;
    slot = hl;                                                     // RocksDB put: the latest value wins
}

// make room: every segment about to be overwritten by the span ending at new_head is copied to pinned host first
static int seglog_spill_for(rafting_engine* e, SegLog* L, uint64_t new_head) {
    const uint64_t last_seg = (new_head - 1) / L->seg_bytes;      // highest logical segment that will hold data
    while (last_seg >= L->nseg && L->spilled_upto <= last_seg - L->nseg) {
        const uint64_t s = L->spilled_upto;
        if (L->cold.size() <= s) { L->cold.resize(s + 1, nullptr); L->cold_ready.resize(s + 1, nullptr); }
        if (s < L->seg_live.size() && L->seg_live[s] == 0) {       // nothing live in it (compacted / overwritten): no cold copy
            L->spilled_upto = s + 1; L->spills_skipped++;
            continue;
        }
        CU(cudaHostAlloc((void**)&L->cold[s], L->seg_bytes, cudaHostAllocDefault));
        CU(cudaEventCreateWithFlags(&L->cold_ready[s], cudaEventDisableTiming));
        CU(cudaEventRecord(e->ev_seg, e->stream));                 // the spill sees every append already enqueued
        CU(cudaStreamWaitEvent(L->s_spill, e->ev_seg, 0));
        CU(cudaMemcpyAsync(L->cold[s], L->arena + (s % L->nseg) * (uint64_t)L->seg_bytes, L->seg_bytes, cudaMemcpyDeviceToHost, L->s_spill));
        CU(cudaEventRecord(L->cold_ready[s], L->s_spill));
        CU(cudaStreamWaitEvent(e->stream, L->cold_ready[s], 0));   // later appends into that arena slot wait for it
        L->spilled_upto = s + 1; L->spilled_bytes += L->seg_bytes;
    }
    return RAFTING_OK;
}

extern "C" int rafting_log_append(rafting_engine_t* e, const rafting_entry_ref_t* refs, uint32_t n, const void* blob, size_t blob_bytes) {
    if (!e || !e->seglog) return fail(RAFTING_E_INVAL, "entry buffer not configured (rafting_log_config)");
    if (n == 0) return RAFTING_OK;
    if (!refs || (!blob && blob_bytes)) return fail(RAFTING_E_INVAL, "null argument");
    SegLog* L = e->seglog;
    CU(cudaSetDevice(e->cfg.device));
    // the payload blob travels to the device once and stays there while the sub-batches below are scattered
    // (pin it on the host for a truly asynchronous copy)
    CU(cudaStreamSynchronize(e->stream));                          // previous append done with d_blob / staging
    if (blob_bytes > L->d_blob_cap) {
        if (L->d_blob) cudaFree(L->d_blob);
        L->d_blob_cap = blob_bytes + blob_bytes / 2 + 256;
        CU(cudaMalloc((void**)&L->d_blob, L->d_blob_cap));
    }
    if (blob_bytes) CU(cudaMemcpyAsync(L->d_blob, blob, blob_bytes, cudaMemcpyHostToDevice, e->stream));
    const uint64_t room = (uint64_t)L->seg_bytes * (L->nseg - 1), arena_bytes = (uint64_t)L->seg_bytes * L->nseg;
    uint32_t done = 0;
    while (done < n) {
        // layout pass (arithmetic only): records never straddle a segment; a sub-batch ends where the arena,
        // minus one segment, would be exceeded
        const size_t need = (size_t)(n - done) * sizeof(AppendRec);
        if (need > L->stage_cap) {
            CU(cudaStreamSynchronize(e->stream));
            if (L->stage) cudaFreeHost(L->stage);
            L->stage_cap = need + need / 2;
            CU(cudaHostAlloc((void**)&L->stage, L->stage_cap, cudaHostAllocDefault));
        } else if (done) CU(cudaStreamSynchronize(e->stream));     // the previous sub-batch has left the staging buffer
        AppendRec* recs = (AppendRec*)L->stage;
        uint64_t head = L->head; uint32_t m = 0;
        for (uint32_t i = done; i < n; i++) {
            if (refs[i].gid >= e->G) return fail(RAFTING_E_INVAL, "ref %u: gid out of range", i);
            if ((uint64_t)refs[i].blob_off + refs[i].len > blob_bytes) return fail(RAFTING_E_INVAL, "ref %u: payload beyond the blob", i);
            const uint64_t rec = sizeof(SegHdr) + (((uint64_t)refs[i].len + 15) & ~15ull);
            if (rec > L->seg_bytes) return fail(RAFTING_E_CAPACITY, "ref %u: record larger than a segment", i);
            uint64_t h2 = head;
            if (h2 / L->seg_bytes != (h2 + rec - 1) / L->seg_bytes) h2 = (h2 / L->seg_bytes + 1) * L->seg_bytes;   // skip the tail
            if (h2 + rec - L->head > room) break;
            AppendRec& r = recs[m];
            r.gid = refs[i].gid; r.len = refs[i].len; r.index = refs[i].index; r.term = refs[i].term; r.loc = h2; r.src = refs[i].blob_off;
            HostLoc hl; hl.off = h2; hl.len = refs[i].len; hl.term = refs[i].term;
            seglog_put(L, refs[i].gid, refs[i].index, hl);
            head = h2 + rec; m++;
        }
        if (m == 0) return fail(RAFTING_E_CAPACITY, "arena too small for a single record");
        int rc = seglog_spill_for(e, L, head); if (rc) return rc;
        const size_t rec_bytes = (size_t)m * sizeof(AppendRec);
        if (rec_bytes > L->d_req_cap) {
            CU(cudaStreamSynchronize(e->stream));
            if (L->d_req) cudaFree(L->d_req);
            L->d_req_cap = rec_bytes * 2;
            CU(cudaMalloc(&L->d_req, L->d_req_cap));
        }
        CU(cudaMemcpyAsync(L->d_req, recs, rec_bytes, cudaMemcpyHostToDevice, e->stream));
        rafting::seglog_scatter_kernel<<<(uint32_t)(((uint64_t)m * 8 + 255) / 256), 256, 0, e->stream>>>(
            (const AppendRec*)L->d_req, m, L->appended, L->d_blob, L->arena, arena_bytes, L->ring, L->K);
        CU(cudaGetLastError());
        L->head = head; L->appended += m; done += m;
    }
    return RAFTING_OK;
}

// stored key range of a group straight from the tables (RocksLog visibility: epoch / truncate are metadata)
static int seglog_range(rafting_engine* e, uint32_t gid, int64_t* lo, int64_t* hi) {
    uint64_t meta;
    CU(cudaMemcpyAsync(&meta, e->T.g_meta + gid, 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(lo, e->T.g_lo + gid, 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(hi, e->T.g_hi + gid, 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if ((((uint32_t)meta >> rafting::W_NRUNS_SH) & 0xf) == 0) { *lo = 1; *hi = 0; }
    return RAFTING_OK;
}
static inline bool seglog_resident(const SegLog* L, uint64_t off) {
    const uint64_t newest = L->head ? (L->head - 1) / L->seg_bytes : 0;
    return off / L->seg_bytes + L->nseg > newest;
}
static int seglog_fetch(rafting_engine* e, SegLog* L, const HostLoc& hl, void* dst) {
    const uint64_t arena_bytes = (uint64_t)L->seg_bytes * L->nseg;
    const uint64_t seg = hl.off / L->seg_bytes;
    if (seglog_resident(L, hl.off)) {                              // still in HBM
        CU(cudaMemcpyAsync(dst, L->arena + (hl.off % arena_bytes) + sizeof(SegHdr), hl.len, cudaMemcpyDeviceToHost, e->stream));
        L->hbm_hits++;
    } else {                                                       // cold tier
        if (seg >= L->cold.size() || !L->cold[seg]) return fail(RAFTING_E_INVAL, "segment %llu neither resident nor spilled", (unsigned long long)seg);
        CU(cudaEventSynchronize(L->cold_ready[seg]));
        memcpy(dst, L->cold[seg] + (hl.off % L->seg_bytes) + sizeof(SegHdr), hl.len);
        L->cold_hits++;
    }
    return RAFTING_OK;
}

extern "C" int rafting_log_read(rafting_engine_t* e, uint32_t gid, int64_t first_index, uint32_t max_n,
                                rafting_entry_ref_t* refs_out, void* blob_out, size_t blob_cap, uint32_t* n_out) {
    if (!e || !e->seglog) return fail(RAFTING_E_INVAL, "entry buffer not configured (rafting_log_config)");
    if (gid >= e->G || !refs_out || !n_out) return fail(RAFTING_E_INVAL, "bad argument");
    SegLog* L = e->seglog;
    CU(cudaSetDevice(e->cfg.device));
    int64_t lo, hi; int rc = seglog_range(e, gid, &lo, &hi); if (rc) return rc;
    uint32_t n = 0; size_t used = 0;
    for (int64_t idx = first_index; n < max_n; idx++) {
        if (idx < lo || idx > hi) break;                           // RocksLog.get: null outside the stored keys
        const HostLoc* hl = seglog_find(L, gid, idx);
        if (!hl) break;
        if (used + hl->len > blob_cap) break;
        if (hl->len) { rc = seglog_fetch(e, L, *hl, (uint8_t*)blob_out + used); if (rc) return rc; }
        refs_out[n].gid = gid; refs_out[n].len = hl->len; refs_out[n].index = idx; refs_out[n].term = hl->term; refs_out[n].blob_off = used;
        used += ((size_t)hl->len + 15) & ~(size_t)15; n++;
    }
    CU(cudaStreamSynchronize(e->stream));
    *n_out = n;
    return RAFTING_OK;
}

// Many ranges at once (the AppendEntries plans of a step): entries resident in HBM are gathered by one
// kernel into a contiguous device buffer and copied down once; the rest come from the cold tier.
// refs_out[k].len == 0xffffffff marks an entry that was never appended.
extern "C" int rafting_log_gather(rafting_engine_t* e, uint32_t n_ranges, const uint32_t* gids, const int64_t* firsts,
                                  const uint32_t* counts, rafting_entry_ref_t* refs_out, uint32_t refs_cap,
                                  void* blob_out, size_t blob_cap, uint32_t* n_out, size_t* bytes_out) {
    if (!e || !e->seglog) return fail(RAFTING_E_INVAL, "entry buffer not configured (rafting_log_config)");
    if (!gids || !firsts || !counts || !refs_out || !n_out) return fail(RAFTING_E_INVAL, "null argument");
    SegLog* L = e->seglog;
    CU(cudaSetDevice(e->cfg.device));
    std::vector<GatherReq> req; std::vector<uint32_t> cold;
    uint32_t m = 0; size_t used = 0;
    for (uint32_t r = 0; r < n_ranges; r++) {
        if (gids[r] >= e->G) return fail(RAFTING_E_INVAL, "range %u: gid out of range", r);
        for (uint32_t k = 0; k < counts[r]; k++) {
            if (m >= refs_cap) return fail(RAFTING_E_CAPACITY, "refs_out too small");
            const int64_t idx = firsts[r] + k;
            rafting_entry_ref_t& o = refs_out[m];
            o.gid = gids[r]; o.index = idx; o.term = 0; o.blob_off = used; o.len = 0xffffffffu;
            const HostLoc* hl = seglog_find(L, gids[r], idx);      // lengths / offsets come from the host index
            if (hl) {
                if (used + hl->len > blob_cap) return fail(RAFTING_E_CAPACITY, "blob_out too small");
                o.len = hl->len; o.term = hl->term;
                if (seglog_resident(L, hl->off)) { GatherReq q; q.gid = gids[r]; q.slot = m; q.index = idx; q.out_off = used; req.push_back(q); }
                else if (hl->len) cold.push_back(m);
                used += ((size_t)hl->len + 15) & ~(size_t)15;
            }
            m++;
        }
    }
    const uint32_t nq = (uint32_t)req.size();
    if (nq) {
        const size_t req_bytes = (size_t)nq * sizeof(GatherReq), len_bytes = (size_t)nq * 4;
        if (req_bytes + len_bytes > L->d_req_cap || used > L->d_out_cap) CU(cudaStreamSynchronize(e->stream));
        if (req_bytes + len_bytes > L->d_req_cap) {
            if (L->d_req) cudaFree(L->d_req);
            L->d_req_cap = (req_bytes + len_bytes) * 2;
            CU(cudaMalloc(&L->d_req, L->d_req_cap));
        }
        if (used > L->d_out_cap) {
            if (L->d_out) cudaFree(L->d_out);
            L->d_out_cap = used * 2;
            CU(cudaMalloc((void**)&L->d_out, L->d_out_cap));
        }
        uint32_t* d_lens = (uint32_t*)((uint8_t*)L->d_req + req_bytes);
        CU(cudaMemcpyAsync(L->d_req, req.data(), req_bytes, cudaMemcpyHostToDevice, e->stream));
        if (!L->t0) { CU(cudaEventCreate(&L->t0)); CU(cudaEventCreate(&L->t1)); }
        const uint64_t newest = L->head ? (L->head - 1) / L->seg_bytes : 0;
        const uint64_t oldest = newest + 1 >= L->nseg ? (newest + 1 - L->nseg) * (uint64_t)L->seg_bytes : 0;
        CU(cudaEventRecord(L->t0, e->stream));
        rafting::seglog_gather_kernel<<<(uint32_t)(((uint64_t)nq * 8 + 255) / 256), 256, 0, e->stream>>>(
            (const GatherReq*)L->d_req, nq, L->ring, L->K, L->arena, (uint64_t)L->seg_bytes * L->nseg, oldest, d_lens, L->d_out);
        CU(cudaGetLastError());
        CU(cudaEventRecord(L->t1, e->stream));
        std::vector<uint32_t> lens(nq);
        CU(cudaMemcpyAsync(lens.data(), d_lens, len_bytes, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(blob_out, L->d_out, used, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        CU(cudaEventElapsedTime(&L->last_gather_kernel_ms, L->t0, L->t1));
        L->last_gather_bytes = used;
        // ring misses (the slot now caches a newer index of the same group): direct copy through the host index
        for (uint32_t k = 0; k < nq; k++) {
            if (lens[k] != 0xffffffffu) { L->hbm_hits++; continue; }
            const rafting_entry_ref_t& o = refs_out[req[k].slot];
            const HostLoc* hl = seglog_find(L, o.gid, o.index);
            if (hl && hl->len) { int rc = seglog_fetch(e, L, *hl, (uint8_t*)blob_out + o.blob_off); if (rc) return rc; }
        }
    }
    for (uint32_t k : cold) {                                      // after the bulk copy-down
        const rafting_entry_ref_t& o = refs_out[k];
        int rc = seglog_fetch(e, L, *seglog_find(L, o.gid, o.index), (uint8_t*)blob_out + o.blob_off); if (rc) return rc;
    }
    CU(cudaStreamSynchronize(e->stream));
    *n_out = m;
    if (bytes_out) *bytes_out = used;
    return RAFTING_OK;
}

// Garbage collection behind RaftLog.flush (RocksLog.java:228-242: deleteRange below the new epoch): index entries below
// a group's lowest stored key are dropped, and a cold (pinned host) segment whose last live record went away is freed.
// A segment with no live record is also never spilled in the first place (seglog_spill_for).
extern "C" int rafting_log_trim(rafting_engine_t* e, uint32_t first_gid, uint32_t count, uint64_t* dropped_entries, uint64_t* freed_cold_bytes) {
    if (!e || !e->seglog) return fail(RAFTING_E_INVAL, "entry buffer not configured (rafting_log_config)");
    if ((uint64_t)first_gid + count > e->G) return fail(RAFTING_E_CAPACITY, "gid range beyond max_groups");
    SegLog* L = e->seglog;
    CU(cudaSetDevice(e->cfg.device));
    std::vector<uint64_t> meta(count); std::vector<int64_t> lo(count), hi(count);
    if (count) {
        CU(cudaMemcpyAsync(meta.data(), e->T.g_meta + first_gid, (size_t)count * 8, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(lo.data(), e->T.g_lo + first_gid, (size_t)count * 8, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(hi.data(), e->T.g_hi + first_gid, (size_t)count * 8, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    uint64_t dropped = 0, freed = 0;
    for (uint32_t k = 0; k < count; k++) {
        rafting::GroupIdx& gi = L->index[first_gid + k];
        if (gi.v.empty()) continue;
        const bool empty = (((uint32_t)meta[k] >> rafting::W_NRUNS_SH) & 0xf) == 0;
        // entries below the lowest stored key are gone for good (an empty store after a flush beyond its end: everything
        // up to the old end); a truncated suffix is NOT dropped here — re-appends overwrite it
        const int64_t keep_from = empty ? hi[k] + 1 : lo[k];
        int64_t cut = keep_from - gi.base;
        if (cut <= 0) continue;
        if (cut > (int64_t)gi.v.size()) cut = (int64_t)gi.v.size();
        for (int64_t i = 0; i < cut; i++) {
            const HostLoc& h = gi.v[(size_t)i];
            if (h.len == 0xffffffffu) continue;
            L->seg_live[h.off / L->seg_bytes]--; L->indexed--; dropped++;
        }
        gi.v.erase(gi.v.begin(), gi.v.begin() + cut);
        gi.base += cut;
    }
    for (size_t s = 0; s < L->cold.size(); s++) {
        if (!L->cold[s] || (s < L->seg_live.size() && L->seg_live[s] != 0)) continue;
        CU(cudaEventSynchronize(L->cold_ready[s]));
        cudaFreeHost(L->cold[s]); L->cold[s] = nullptr;
        cudaEventDestroy(L->cold_ready[s]); L->cold_ready[s] = nullptr;
        freed += L->seg_bytes;
    }
    L->trimmed += dropped; L->cold_freed_bytes += freed;
    if (dropped_entries) *dropped_entries = dropped;
    if (freed_cold_bytes) *freed_cold_bytes = freed;
    return RAFTING_OK;
}

extern "C" int rafting_log_stats(rafting_engine_t* e, uint64_t* out, uint32_t n) {
    if (!e || !e->seglog || !out) return fail(RAFTING_E_INVAL, "entry buffer not configured");
    SegLog* L = e->seglog;
    const uint64_t v[] = {L->appended, L->head, L->spilled_bytes, L->hbm_hits, L->cold_hits, L->indexed,
                          (uint64_t)(L->last_gather_kernel_ms * 1e6), L->last_gather_bytes,
                          L->trimmed, L->cold_freed_bytes, L->spills_skipped};
    for (uint32_t i = 0; i < n && i < 11; i++) out[i] = v[i];
    return 11;
}
