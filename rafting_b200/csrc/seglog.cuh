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
//   RocksLog.truncate / flush   (deleteRange)  -> nothing to do here: visibility of a record is decided by
//                                                 the group's stored key range [g_lo, g_hi] in the tables,
//                                                 and a re-appended index overwrites its ring slot
//
// Record layout inside a segment: 24-byte header {gid u32, len u32, index i64, term i64} + payload padded
// to 8 bytes.  Records never straddle segments.  A logical byte offset (segment number * seg_bytes + offset)
// names a record forever; the HBM arena holds the newest `nseg` segments, older ones are read from the
// host copy.  Per group a device ring of K slots maps index -> logical offset for the gather kernel.
//
// Included at the end of engine.cu (same translation unit: it uses rafting_engine, fail(), CU()).
#pragma once
#include <unordered_map>

namespace rafting {

struct SegHdr { uint32_t gid, len; int64_t index, term; };
static_assert(sizeof(SegHdr) == 24, "segment record header");

struct HostLoc { uint64_t off; uint32_t len; int64_t term; };

struct SegLog {
    uint32_t seg_bytes = 0, nseg = 0, K = 0;
    uint8_t* arena = nullptr;                 // [nseg * seg_bytes] in HBM
    uint64_t head = 0;                        // logical offset of the next record
    uint64_t spilled_upto = 0;                // logical segments < this have a host copy enqueued
    int64_t*  ring_index = nullptr;           // [G * K]
    uint64_t* ring_loc = nullptr;             // [G * K] logical offset of the record header
    uint32_t* ring_len = nullptr;             // [G * K]
    std::vector<uint8_t*> cold;               // pinned host copy per logical segment (null until spilled)
    std::vector<cudaEvent_t> cold_ready;      // spill completion per logical segment
    std::unordered_map<uint64_t, HostLoc> index;   // (gid, index) -> record; authoritative for lengths and the cold tier
    uint8_t* stage = nullptr; size_t stage_cap = 0;       // pinned staging of one append batch
    void* d_req = nullptr; size_t d_req_cap = 0;           // device scratch of the gather requests
    uint8_t* d_out = nullptr; size_t d_out_cap = 0;
    cudaStream_t s_spill = nullptr;
    uint64_t appended = 0, spilled_bytes = 0, hbm_hits = 0, cold_hits = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;    // device time of the last gather's kernels
    float last_gather_kernel_ms = 0; uint64_t last_gather_bytes = 0;
};

static inline uint64_t seg_key(uint32_t gid, int64_t index) { return ((uint64_t)gid << 40) ^ (uint64_t)index; }

// one thread per appended record: publish (index -> logical offset) in the group's ring
__global__ void seglog_index_kernel(const SegHdr* __restrict__ hdrs, const uint64_t* __restrict__ locs, uint32_t n,
                                    int64_t* ring_index, uint64_t* ring_loc, uint32_t* ring_len, uint32_t K, uint32_t G) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const SegHdr h = hdrs[t];
    const uint64_t loc = locs[t];
    if (h.gid >= G || (loc >> 63)) return;             // bit 63: a later record of this batch owns the slot
    const size_t s = (size_t)h.gid * K + ((uint64_t)h.index & (K - 1));
    ring_index[s] = h.index; ring_loc[s] = loc; ring_len[s] = h.len;
}

struct GatherReq { uint32_t gid; uint32_t slot; int64_t index; uint64_t out_off; };   // slot: position in the request list

// pass 1: one thread per requested entry -> its length (0xffffffff when the ring does not hold it in HBM)
__global__ void seglog_probe_kernel(const GatherReq* __restrict__ req, uint32_t n, const int64_t* __restrict__ ring_index,
                                    const uint64_t* __restrict__ ring_loc, const uint32_t* __restrict__ ring_len,
                                    uint32_t K, uint64_t oldest_resident, uint32_t* __restrict__ lens) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const GatherReq q = req[t];
    const size_t s = (size_t)q.gid * K + ((uint64_t)q.index & (K - 1));
    const bool hit = ring_index[s] == q.index && ring_loc[s] >= oldest_resident;
    lens[t] = hit ? ring_len[s] : 0xffffffffu;
}
// pass 2: one warp per entry, 16-byte vector copies HBM arena -> contiguous output (payload only)
__global__ void seglog_copy_kernel(const GatherReq* __restrict__ req, uint32_t n, const uint64_t* __restrict__ ring_loc,
                                   const uint32_t* __restrict__ ring_len, uint32_t K, const uint8_t* __restrict__ arena,
                                   uint64_t arena_bytes, const uint32_t* __restrict__ lens, uint8_t* __restrict__ out) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
    if (w >= n) return;
    if (lens[w] == 0xffffffffu) return;
    const GatherReq q = req[w];
    const size_t s = (size_t)q.gid * K + ((uint64_t)q.index & (K - 1));
    const uint8_t* src = arena + (ring_loc[s] % arena_bytes) + sizeof(SegHdr);      // records are 8-byte aligned
    uint8_t* dst = out + q.out_off;                                                  // out_off is 8-byte aligned
    const uint32_t len = ring_len[s], words = (len + 7) / 8;
    for (uint32_t i = lane; i < words; i += 32) ((uint64_t*)dst)[i] = ((const uint64_t*)src)[i];
}

}  // namespace rafting

using rafting::SegLog; using rafting::SegHdr; using rafting::HostLoc; using rafting::GatherReq;

static void seglog_release(rafting_engine* e) {
    SegLog* L = e->seglog; if (!L) return;
    if (L->s_spill) { cudaStreamSynchronize(L->s_spill); cudaStreamDestroy(L->s_spill); }
    for (auto ev : L->cold_ready) if (ev) cudaEventDestroy(ev);
    for (auto p : L->cold) if (p) cudaFreeHost(p);
    if (L->arena) cudaFree(L->arena);
    if (L->ring_index) cudaFree(L->ring_index);
    if (L->ring_loc) cudaFree(L->ring_loc);
    if (L->ring_len) cudaFree(L->ring_len);
    if (L->stage) cudaFreeHost(L->stage);
    if (L->d_req) cudaFree(L->d_req);
    if (L->d_out) cudaFree(L->d_out);
    if (L->t0) { cudaEventDestroy(L->t0); cudaEventDestroy(L->t1); }
    delete L; e->seglog = nullptr;
}

extern "C" int rafting_log_config(rafting_engine_t* e, uint32_t segment_bytes, uint32_t hbm_segments, uint32_t ring_slots) {
    if (!e) return fail(RAFTING_E_INVAL, "null argument");
    if (e->seglog) return fail(RAFTING_E_INVAL, "entry buffer already configured");
    if (segment_bytes < 4096 || (segment_bytes & 7) || hbm_segments < 2 || ring_slots == 0 || (ring_slots & (ring_slots - 1)))
        return fail(RAFTING_E_INVAL, "segment_bytes >= 4096 and 8-aligned, hbm_segments >= 2, ring_slots a power of two");
    CU(cudaSetDevice(e->cfg.device));
    SegLog* L = new SegLog();
    L->seg_bytes = segment_bytes; L->nseg = hbm_segments; L->K = ring_slots;
    e->seglog = L;
    const size_t ring = (size_t)e->G * ring_slots;
    CU(cudaMalloc(&L->arena, (size_t)segment_bytes * hbm_segments));
    CU(cudaMalloc(&L->ring_index, ring * 8)); CU(cudaMalloc(&L->ring_loc, ring * 8)); CU(cudaMalloc(&L->ring_len, ring * 4));
    CU(cudaMemset(L->ring_index, 0xff, ring * 8));                 // index -1 never matches
    CU(cudaMemset(L->ring_loc, 0, ring * 8)); CU(cudaMemset(L->ring_len, 0, ring * 4));
    CU(cudaStreamCreateWithFlags(&L->s_spill, cudaStreamNonBlocking));
    return RAFTING_OK;
}

// make room: every segment about to be overwritten by [head, head + need) is copied to pinned host first
static int seglog_spill_for(rafting_engine* e, SegLog* L, uint64_t new_head) {
    const uint64_t last_seg = (new_head - 1) / L->seg_bytes;             // highest logical segment that will hold data
    // segment s may live in the arena while s > last_seg - nseg
    while (last_seg >= L->nseg && L->spilled_upto <= last_seg - L->nseg) {
        const uint64_t s = L->spilled_upto;
        if (L->cold.size() <= s) { L->cold.resize(s + 1, nullptr); L->cold_ready.resize(s + 1, nullptr); }
        CU(cudaHostAlloc((void**)&L->cold[s], L->seg_bytes, cudaHostAllocDefault));
        CU(cudaEventCreateWithFlags(&L->cold_ready[s], cudaEventDisableTiming));
        // the spill must see every append already enqueued on the engine stream
        CU(cudaEventRecord(e->ev_seg, e->stream));
        CU(cudaStreamWaitEvent(L->s_spill, e->ev_seg, 0));
        CU(cudaMemcpyAsync(L->cold[s], L->arena + (s % L->nseg) * (uint64_t)L->seg_bytes, L->seg_bytes, cudaMemcpyDeviceToHost, L->s_spill));
        CU(cudaEventRecord(L->cold_ready[s], L->s_spill));
        // ... and later appends into that arena slot must wait for the spill
        CU(cudaStreamWaitEvent(e->stream, L->cold_ready[s], 0));
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
    if (!e->ev_seg) CU(cudaEventCreateWithFlags(&e->ev_seg, cudaEventDisableTiming));
    // layout pass: records never straddle a segment; a batch that would not fit the arena is cut and the
    // remainder appended by a second pass
    std::vector<uint64_t> locs(n);
    uint64_t head = L->head, total = 0;
    const uint64_t room = (uint64_t)L->seg_bytes * (L->nseg - 1);
    const uint32_t n_all = n;
    for (uint32_t i = 0; i < n; i++) {
        if (refs[i].gid >= e->G) return fail(RAFTING_E_INVAL, "ref %u: gid out of range", i);
        if ((uint64_t)refs[i].blob_off + refs[i].len > blob_bytes) return fail(RAFTING_E_INVAL, "ref %u: payload beyond the blob", i);
        const uint64_t rec = sizeof(SegHdr) + (((uint64_t)refs[i].len + 7) & ~7ull);
        if (rec > L->seg_bytes) return fail(RAFTING_E_CAPACITY, "ref %u: record larger than a segment", i);
        uint64_t h2 = head;
        if (h2 / L->seg_bytes != (h2 + rec - 1) / L->seg_bytes) h2 = (h2 / L->seg_bytes + 1) * L->seg_bytes;   // skip the tail
        if (h2 + rec - L->head > room) { n = i; break; }
        locs[i] = h2; head = h2 + rec;
    }
    if (n == 0) return fail(RAFTING_E_CAPACITY, "arena too small for a single record batch");
    total = head - L->head;
    // staging image = [span of bytes head..new head) + headers array + locs array
    const size_t hdr_bytes = (size_t)n * sizeof(SegHdr), loc_bytes = (size_t)n * 8;
    const size_t need = total + hdr_bytes + loc_bytes + 64;
    if (need > L->stage_cap) {
        CU(cudaStreamSynchronize(e->stream));
        if (L->stage) cudaFreeHost(L->stage);
        if (L->d_req) { cudaFree(L->d_req); L->d_req = nullptr; L->d_req_cap = 0; }
        L->stage_cap = need + need / 2;
        CU(cudaHostAlloc((void**)&L->stage, L->stage_cap, cudaHostAllocDefault));
    } else CU(cudaStreamSynchronize(e->stream));                      // the previous batch has left the staging buffer
    memset(L->stage, 0, total);
    std::unordered_map<uint64_t, uint32_t> slot_owner;
    SegHdr* hdrs = (SegHdr*)(L->stage + total);
    uint64_t* hlocs = (uint64_t*)(L->stage + total + hdr_bytes);
    for (uint32_t i = 0; i < n; i++) {
        SegHdr h; h.gid = refs[i].gid; h.len = refs[i].len; h.index = refs[i].index; h.term = refs[i].term;
        uint8_t* dst = L->stage + (locs[i] - L->head);
        memcpy(dst, &h, sizeof(h));
        if (refs[i].len) memcpy(dst + sizeof(h), (const uint8_t*)blob + refs[i].blob_off, refs[i].len);
        hdrs[i] = h; hlocs[i] = locs[i];
        slot_owner[((uint64_t)refs[i].gid << 32) | (uint32_t)((uint64_t)refs[i].index & (L->K - 1))] = i;   // RocksDB put order: the last one wins
        HostLoc hl; hl.off = locs[i]; hl.len = refs[i].len; hl.term = refs[i].term;
        L->index[rafting::seg_key(refs[i].gid, refs[i].index)] = hl;              // RocksDB put: the latest value wins
    }
    for (uint32_t i = 0; i < n; i++)
        if (slot_owner[((uint64_t)refs[i].gid << 32) | (uint32_t)((uint64_t)refs[i].index & (L->K - 1))] != i) hlocs[i] |= 1ull << 63;
    int rc = seglog_spill_for(e, L, head); if (rc) return rc;
    // H2D of the span (split where it wraps around the arena)
    const uint64_t arena_bytes = (uint64_t)L->seg_bytes * L->nseg;
    uint64_t off = L->head, left = total, src = 0;
    while (left) {
        const uint64_t a = off % arena_bytes, chunk = left < arena_bytes - a ? left : arena_bytes - a;
        CU(cudaMemcpyAsync(L->arena + a, L->stage + src, chunk, cudaMemcpyHostToDevice, e->stream));
        off += chunk; src += chunk; left -= chunk;
    }
    // headers + locs for the index kernel
    if (hdr_bytes + loc_bytes > L->d_req_cap) {
        if (L->d_req) cudaFree(L->d_req);
        L->d_req_cap = (hdr_bytes + loc_bytes) * 2;
        CU(cudaMalloc(&L->d_req, L->d_req_cap));
    }
    CU(cudaMemcpyAsync(L->d_req, hdrs, hdr_bytes + loc_bytes, cudaMemcpyHostToDevice, e->stream));
    rafting::seglog_index_kernel<<<(n + 255) / 256, 256, 0, e->stream>>>((const SegHdr*)L->d_req, (const uint64_t*)((uint8_t*)L->d_req + hdr_bytes), n,
                                                                         L->ring_index, L->ring_loc, L->ring_len, L->K, e->G);
    CU(cudaGetLastError());
    L->head = head; L->appended += n;
    if (n < n_all) return rafting_log_append(e, refs + n, n_all - n, blob, blob_bytes);
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
static int seglog_fetch(rafting_engine* e, SegLog* L, const HostLoc& hl, void* dst) {
    const uint64_t arena_bytes = (uint64_t)L->seg_bytes * L->nseg;
    const uint64_t seg = hl.off / L->seg_bytes;
    const uint64_t newest = L->head ? (L->head - 1) / L->seg_bytes : 0;
    if (seg + L->nseg > newest) {                                                  // still resident in HBM
        CU(cudaMemcpyAsync(dst, L->arena + (hl.off % arena_bytes) + sizeof(SegHdr), hl.len, cudaMemcpyDeviceToHost, e->stream));
        L->hbm_hits++;
    } else {                                                                       // cold tier
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
        if (idx < lo || idx > hi) break;                                           // RocksLog.get: null outside the stored keys
        auto it = L->index.find(rafting::seg_key(gid, idx));
        if (it == L->index.end()) break;
        const HostLoc& hl = it->second;
        if (used + hl.len > blob_cap) break;
        if (hl.len) { rc = seglog_fetch(e, L, hl, (uint8_t*)blob_out + used); if (rc) return rc; }
        refs_out[n].gid = gid; refs_out[n].len = hl.len; refs_out[n].index = idx; refs_out[n].term = hl.term; refs_out[n].blob_off = used;
        used += ((size_t)hl.len + 7) & ~(size_t)7; n++;
    }
    CU(cudaStreamSynchronize(e->stream));
    *n_out = n;
    return RAFTING_OK;
}

// Many ranges at once (the AppendEntries plans of a step): entries resident in HBM are gathered by a
// kernel into one contiguous device buffer and copied down once; the rest come from the cold tier.
// refs_out[k].len == 0xffffffff marks an entry that is not stored (beyond the key range or never appended).
extern "C" int rafting_log_gather(rafting_engine_t* e, uint32_t n_ranges, const uint32_t* gids, const int64_t* firsts,
                                  const uint32_t* counts, rafting_entry_ref_t* refs_out, uint32_t refs_cap,
                                  void* blob_out, size_t blob_cap, uint32_t* n_out, size_t* bytes_out) {
    if (!e || !e->seglog) return fail(RAFTING_E_INVAL, "entry buffer not configured (rafting_log_config)");
    if (!gids || !firsts || !counts || !refs_out || !n_out) return fail(RAFTING_E_INVAL, "null argument");
    SegLog* L = e->seglog;
    CU(cudaSetDevice(e->cfg.device));
    // expand the ranges; lengths and output offsets come from the host index (authoritative)
    std::vector<GatherReq> req; std::vector<uint32_t> cold;
    uint32_t m = 0; size_t used = 0;
    const uint64_t newest = L->head ? (L->head - 1) / L->seg_bytes : 0;
    for (uint32_t r = 0; r < n_ranges; r++) {
        if (gids[r] >= e->G) return fail(RAFTING_E_INVAL, "range %u: gid out of range", r);
        for (uint32_t k = 0; k < counts[r]; k++) {
            if (m >= refs_cap) return fail(RAFTING_E_CAPACITY, "refs_out too small");
            const int64_t idx = firsts[r] + k;
            rafting_entry_ref_t& o = refs_out[m];
            o.gid = gids[r]; o.index = idx; o.term = 0; o.blob_off = used; o.len = 0xffffffffu;
            auto it = L->index.find(rafting::seg_key(gids[r], idx));
            if (it != L->index.end()) {
                const HostLoc& hl = it->second;
                if (used + hl.len > blob_cap) return fail(RAFTING_E_CAPACITY, "blob_out too small");
                o.len = hl.len; o.term = hl.term;
                const bool resident = hl.off / L->seg_bytes + L->nseg > newest;
                if (resident) { GatherReq q; q.gid = gids[r]; q.slot = m; q.index = idx; q.out_off = used; req.push_back(q); }
                else if (hl.len) cold.push_back(m);
                used += ((size_t)hl.len + 7) & ~(size_t)7;
            }
            m++;
        }
    }
    const uint32_t nq = (uint32_t)req.size();
    if (nq) {
        const size_t req_bytes = (size_t)nq * sizeof(GatherReq), len_bytes = (size_t)nq * 4;
        if (req_bytes + len_bytes > L->d_req_cap) {
            CU(cudaStreamSynchronize(e->stream));
            if (L->d_req) cudaFree(L->d_req);
            L->d_req_cap = (req_bytes + len_bytes) * 2;
            CU(cudaMalloc(&L->d_req, L->d_req_cap));
        }
        if (used > L->d_out_cap) {
            CU(cudaStreamSynchronize(e->stream));
            if (L->d_out) cudaFree(L->d_out);
            L->d_out_cap = used * 2;
            CU(cudaMalloc((void**)&L->d_out, L->d_out_cap));
        }
        uint32_t* d_lens = (uint32_t*)((uint8_t*)L->d_req + req_bytes);
        CU(cudaMemcpyAsync(L->d_req, req.data(), req_bytes, cudaMemcpyHostToDevice, e->stream));
        if (!L->t0) { CU(cudaEventCreate(&L->t0)); CU(cudaEventCreate(&L->t1)); }
        CU(cudaEventRecord(L->t0, e->stream));
        const uint64_t oldest = newest + 1 >= L->nseg ? (newest + 1 - L->nseg) * (uint64_t)L->seg_bytes : 0;
        rafting::seglog_probe_kernel<<<(nq + 255) / 256, 256, 0, e->stream>>>((const GatherReq*)L->d_req, nq, L->ring_index, L->ring_loc,
                                                                               L->ring_len, L->K, oldest, d_lens);
        rafting::seglog_copy_kernel<<<(nq * 32 + 255) / 256, 256, 0, e->stream>>>((const GatherReq*)L->d_req, nq, L->ring_loc, L->ring_len, L->K,
                                                                                  L->arena, (uint64_t)L->seg_bytes * L->nseg, d_lens, L->d_out);
        CU(cudaGetLastError());
        CU(cudaEventRecord(L->t1, e->stream));
        std::vector<uint32_t> lens(nq);
        CU(cudaMemcpyAsync(lens.data(), d_lens, len_bytes, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(blob_out, L->d_out, used, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        CU(cudaEventElapsedTime(&L->last_gather_kernel_ms, L->t0, L->t1));
        L->last_gather_bytes = used;
        // ring misses (the slot now belongs to a newer index of the same group): direct copy from the arena
        for (uint32_t k = 0; k < nq; k++) {
            if (lens[k] != 0xffffffffu) { L->hbm_hits++; continue; }
            const rafting_entry_ref_t& o = refs_out[req[k].slot];
            auto it = L->index.find(rafting::seg_key(o.gid, o.index));
            if (it != L->index.end() && it->second.len) { int rc = seglog_fetch(e, L, it->second, (uint8_t*)blob_out + o.blob_off); if (rc) return rc; }
        }
    }
    for (uint32_t k : cold) {                                                       // after the bulk copy-down
        const rafting_entry_ref_t& o = refs_out[k];
        auto it = L->index.find(rafting::seg_key(o.gid, o.index));
        int rc = seglog_fetch(e, L, it->second, (uint8_t*)blob_out + o.blob_off); if (rc) return rc;
    }
    CU(cudaStreamSynchronize(e->stream));
    *n_out = m;
    if (bytes_out) *bytes_out = used;
    return RAFTING_OK;
}

extern "C" int rafting_log_stats(rafting_engine_t* e, uint64_t* out, uint32_t n) {
    if (!e || !e->seglog || !out) return fail(RAFTING_E_INVAL, "entry buffer not configured");
    SegLog* L = e->seglog;
    const uint64_t v[] = {L->appended, L->head, L->spilled_bytes, L->hbm_hits, L->cold_hits, (uint64_t)L->index.size(),
                          (uint64_t)(L->last_gather_kernel_ms * 1e6), L->last_gather_bytes};
    for (uint32_t i = 0; i < n && i < 8; i++) out[i] = v[i];
    return 8;
}
