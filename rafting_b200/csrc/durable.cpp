// durable.cpp — write-ahead journal of (term, votedFor) / milestone records, one fdatasync per engine step.
// Host-only (g++), see include/rafting_durable.h for what it replaces in the reference and the file formats.
#include "../../include/rafting_durable.h"

#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

namespace {

enum { OK = 0, E_INVAL = -1, E_NOMEM = -2, E_IO = -3, E_CLOSED = -4, E_CAPACITY = -5 };
thread_local char g_err[512];
int fail(int rc, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return rc;
}

// CRC-32C (Castagnoli), table driven
uint32_t crc_table[256];
bool crc_ready = false;
void crc_init() {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
        crc_table[i] = c;
    }
    crc_ready = true;
}
uint32_t crc32c(const void* p, size_t n, uint32_t c = 0) {
    if (!crc_ready) crc_init();
    const uint8_t* b = (const uint8_t*)p;
    c = ~c;
    for (size_t i = 0; i < n; i++) c = crc_table[(c ^ b[i]) & 0xff] ^ (c >> 8);
    return ~c;
}

struct Row { int64_t ms_index, ms_term, term; int32_t ballot; uint32_t flags; };          // stable.tbl, 32 bytes
static_assert(sizeof(Row) == 32, "table row");
enum { KIND_VOTE = 1, KIND_MILESTONE = 2 };
struct Rec { uint32_t gid; uint16_t kind; int16_t ballot; int64_t a, b; };                 // stable.wal record, 24 bytes
static_assert(sizeof(Rec) == 24, "journal record");
struct Hdr { uint32_t magic, n; uint64_t seq; uint32_t crc, _pad; };                       // batch header, 24 bytes
static_assert(sizeof(Hdr) == 24, "batch header");
constexpr uint32_t MAGIC = 0x52464A31u;   // "RFJ1"

bool write_all(int fd, const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    while (n) {
        ssize_t w = write(fd, b, n);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        b += w; n -= (size_t)w;
    }
    return true;
}
void put_be64(uint8_t* p, int64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)((uint64_t)v >> (56 - 8 * i)); }
void put_be32(uint8_t* p, uint32_t v) { for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (24 - 8 * i)); }

}  // namespace

struct rafting_journal {
    std::string dir;
    uint32_t G = 0;
    int fd_tbl = -1, fd_wal = -1;
    std::vector<Row> rows;          // the recovered + applied state (what restore() answers from)
    std::vector<uint8_t> buf;       // batch under construction
    uint64_t seq = 0, batches = 0, records = 0, syncs = 0, wal_bytes = 0;
    bool failed = false;            // sticky: a batch could not be written AND could not be rolled back
};

static void apply(rafting_journal* j, const Rec& r) {
    Row& w = j->rows[r.gid];
    if (r.kind == KIND_VOTE) { w.term = r.a; w.ballot = r.ballot; }
    else if (r.kind == KIND_MILESTONE) { w.ms_index = r.a; w.ms_term = r.b; }
}

// A batch is either durable as a whole or absent: when the write or the barrier fails, the file is cut back to the
// pre-batch offset and the sequence number is not consumed, so a later batch can never be appended behind torn bytes
// (recovery stops at the first torn batch and would silently drop everything after it — acknowledged votes included).
// If the cut-back itself fails the journal turns sticky-failed: every later commit answers E_IO until it is reopened.
static int commit_buf(rafting_journal* j, uint32_t n) {
    if (j->failed) return fail(E_IO, "journal is failed (an earlier batch could not be rolled back): reopen it");
    Hdr* h = (Hdr*)j->buf.data();
    h->magic = MAGIC; h->n = n; h->seq = j->seq + 1; h->_pad = 0;
    h->crc = crc32c(j->buf.data() + sizeof(Hdr), (size_t)n * sizeof(Rec), crc32c(&h->seq, 8));
    const char* what = nullptr; int err = 0;
    if (!write_all(j->fd_wal, j->buf.data(), j->buf.size())) { what = "journal write"; err = errno; }
    else if (fdatasync(j->fd_wal) != 0) { what = "fdatasync"; err = errno; }
    if (what) {
        const off_t back = (off_t)j->wal_bytes;
        if (ftruncate(j->fd_wal, back) != 0 || lseek(j->fd_wal, back, SEEK_SET) < 0 || fdatasync(j->fd_wal) != 0) j->failed = true;
        return fail(E_IO, "%s failed: %s%s", what, strerror(err), j->failed ? " (and the batch could not be rolled back: journal failed)" : " (batch rolled back)");
    }
    j->seq = h->seq;
    j->syncs++; j->batches++; j->records += n; j->wal_bytes += j->buf.size();
    const Rec* r = (const Rec*)(j->buf.data() + sizeof(Hdr));
    for (uint32_t k = 0; k < n; k++) apply(j, r[k]);
    return OK;
}

extern "C" const char* rafting_durable_last_error(void) { return g_err; }

extern "C" int rafting_journal_open(const char* dir, uint32_t max_groups, rafting_journal_t** out) {
    if (!dir || !out || max_groups == 0) return fail(E_INVAL, "bad argument");
    if (mkdir(dir, 0755) != 0 && errno != EEXIST) return fail(E_IO, "mkdir %s: %s", dir, strerror(errno));
    rafting_journal* j = new rafting_journal();
    j->dir = dir; j->G = max_groups;
    j->rows.assign(max_groups, Row{0, 0, 0, -1, 0});
    const std::string tbl = j->dir + "/stable.tbl", wal = j->dir + "/stable.wal";
    j->fd_tbl = open(tbl.c_str(), O_RDWR | O_CREAT, 0644);
    j->fd_wal = open(wal.c_str(), O_RDWR | O_CREAT, 0644);
    if (j->fd_tbl < 0 || j->fd_wal < 0) { int rc = fail(E_IO, "open journal files in %s: %s", dir, strerror(errno)); rafting_journal_close(j); return rc; }
    // the directory entries of the two files must survive a crash too: one fsync of the directory after creating them
    {
        const int dfd = open(dir, O_RDONLY | O_DIRECTORY);
        if (dfd < 0 || fsync(dfd) != 0) {
            int rc = fail(E_IO, "fsync of directory %s: %s", dir, strerror(errno));
            if (dfd >= 0) close(dfd);
            rafting_journal_close(j); return rc;
        }
        close(dfd);
    }
    // table: as many complete rows as the file holds (a fresh file holds none)
    struct stat sb;
    if (fstat(j->fd_tbl, &sb) == 0 && sb.st_size > 0) {
        size_t have = (size_t)sb.st_size / sizeof(Row);
        if (have > max_groups) have = max_groups;
        if (pread(j->fd_tbl, j->rows.data(), have * sizeof(Row), 0) != (ssize_t)(have * sizeof(Row))) {
            int rc = fail(E_IO, "short read of stable.tbl"); rafting_journal_close(j); return rc;
        }
    }
    // journal: every complete batch, in order; the first torn / foreign one ends the replay and is cut off
    off_t pos = 0;
    for (;;) {
        Hdr h;
        if (pread(j->fd_wal, &h, sizeof(h), pos) != (ssize_t)sizeof(h)) break;
        if (h.magic != MAGIC || h.n > (1u << 26) || h.seq != j->seq + 1) break;
        std::vector<Rec> recs(h.n);
        const size_t bytes = (size_t)h.n * sizeof(Rec);
        if (pread(j->fd_wal, recs.data(), bytes, pos + (off_t)sizeof(h)) != (ssize_t)bytes) break;
        if (crc32c(recs.data(), bytes, crc32c(&h.seq, 8)) != h.crc) break;
        bool ok = true;
        for (const Rec& r : recs) if (r.gid >= max_groups) ok = false;
        if (!ok) break;
        for (const Rec& r : recs) apply(j, r);
        j->seq = h.seq;
        pos += (off_t)(sizeof(h) + bytes);
    }
    if (ftruncate(j->fd_wal, pos) != 0 || lseek(j->fd_wal, pos, SEEK_SET) < 0) {
        int rc = fail(E_IO, "cannot position the journal: %s", strerror(errno)); rafting_journal_close(j); return rc;
    }
    j->wal_bytes = (uint64_t)pos;
    *out = j;
    return OK;
}

extern "C" int rafting_journal_close(rafting_journal_t* j) {
    if (!j) return OK;
    if (j->fd_tbl >= 0) close(j->fd_tbl);
    if (j->fd_wal >= 0) close(j->fd_wal);
    delete j;
    return OK;
}

extern "C" int rafting_journal_commit_step(rafting_journal_t* j, const uint32_t* gids, uint32_t n, int compact,
                                           const uint32_t* role_word, const int64_t* current_term, uint64_t* n_records) {
    if (!j || !role_word || !current_term) return fail(E_INVAL, "null argument");
    if (!gids && n > j->G) return fail(E_CAPACITY, "n > max_groups");
    j->buf.resize(sizeof(Hdr));
    uint32_t cnt = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t gid = gids ? gids[i] : i;
        if (gid >= j->G) return fail(E_CAPACITY, "gid %u beyond max_groups", gid);
        const uint32_t col = (gids && !compact) ? gid : i;
        const uint32_t w = role_word[col];
        if (!(w & (1u << 30))) continue;                                   // persist-dirty: a role object was constructed
        Rec r; r.gid = gid; r.kind = KIND_VOTE; r.ballot = (int16_t)((int)((w >> 8) & 0xff) - 1); r.a = current_term[col]; r.b = 0;
        const size_t at = j->buf.size();
        j->buf.resize(at + sizeof(Rec));
        memcpy(j->buf.data() + at, &r, sizeof(Rec));
        cnt++;
    }
    if (n_records) *n_records = cnt;
    if (cnt == 0) return OK;
    return commit_buf(j, cnt);
}

extern "C" int rafting_journal_milestone(rafting_journal_t* j, uint32_t gid, int64_t index, int64_t term) {
    if (!j) return fail(E_INVAL, "null argument");
    if (gid >= j->G) return fail(E_CAPACITY, "gid beyond max_groups");
    j->buf.resize(sizeof(Hdr) + sizeof(Rec));
    Rec r; r.gid = gid; r.kind = KIND_MILESTONE; r.ballot = 0; r.a = index; r.b = term;
    memcpy(j->buf.data() + sizeof(Hdr), &r, sizeof(Rec));
    return commit_buf(j, 1);
}

extern "C" int rafting_journal_restore(rafting_journal_t* j, uint32_t gid, rafting_stable_t* out) {
    if (!j || !out) return fail(E_INVAL, "null argument");
    if (gid >= j->G) return fail(E_CAPACITY, "gid beyond max_groups");
    const Row& w = j->rows[gid];
    out->term = w.term; out->ballot = w.ballot; out->_pad = 0; out->milestone_index = w.ms_index; out->milestone_term = w.ms_term;
    return OK;
}

extern "C" int rafting_journal_checkpoint(rafting_journal_t* j) {
    if (!j) return fail(E_INVAL, "null argument");
    if (j->failed) return fail(E_IO, "journal is failed: reopen it");
    // 1. the table becomes durable first; only then may the journal that produced it disappear
    if (pwrite(j->fd_tbl, j->rows.data(), j->rows.size() * sizeof(Row), 0) != (ssize_t)(j->rows.size() * sizeof(Row)))
        return fail(E_IO, "table write failed: %s", strerror(errno));
    if (fsync(j->fd_tbl) != 0) return fail(E_IO, "fsync(table) failed: %s", strerror(errno));
    j->syncs++;
    if (ftruncate(j->fd_wal, 0) != 0 || lseek(j->fd_wal, 0, SEEK_SET) < 0) return fail(E_IO, "journal truncate failed: %s", strerror(errno));
    if (fsync(j->fd_wal) != 0) return fail(E_IO, "fsync(journal) failed: %s", strerror(errno));
    j->syncs++;
    j->seq = 0; j->wal_bytes = 0;
    return OK;
}

extern "C" int rafting_journal_stats(rafting_journal_t* j, uint64_t out[4]) {
    if (!j || !out) return fail(E_INVAL, "null argument");
    out[0] = j->batches; out[1] = j->records; out[2] = j->syncs; out[3] = j->wal_bytes;
    return OK;
}

extern "C" int rafting_stable_image(rafting_journal_t* j, uint32_t gid, const void* id_bytes, uint32_t id_len,
                                    void* out, uint32_t cap, uint32_t* len) {
    if (!j || !out || !len) return fail(E_INVAL, "null argument");
    if (gid >= j->G) return fail(E_CAPACITY, "gid beyond max_groups");
    const Row& w = j->rows[gid];
    const uint32_t idn = (w.ballot >= 0 && id_bytes) ? id_len : 0;       // null ballot: length 0, no bytes (StableLock.java:58-63)
    if (cap < 28 + idn) return fail(E_CAPACITY, "image needs %u bytes", 28 + idn);
    uint8_t* p = (uint8_t*)out;
    put_be64(p, w.ms_index); put_be64(p + 8, w.ms_term); put_be64(p + 16, w.term); put_be32(p + 24, idn);
    if (idn) memcpy(p + 28, id_bytes, idn);
    *len = 28 + idn;
    return OK;
}
