// The sharded loop query between PROCESSES with nothing but the C ABI and POSIX: no torch, no NCCL, no MPI, and no
// CUDA runtime in this program.  The parent forks one child per GPU (before anything touches CUDA); each child builds its
// shard of a synthetic keyframe database, exports the CUDA-IPC handle of its exchange buffer (aslam_loopdb_shard_export),
// the handles travel through pipes, aslam_loopdb_shard_init opens the peers' buffers, and every query is ONE kernel
// launch per process whose tail exchanges the keys over NVLink (aslam_loopdb_query_global_sharded_device).  Rank 0 also
// holds the whole database on its GPU: the sharded answer must equal it bit for bit, on every rank, query after query.
//   shard_ipc_selftest [ranks=2] [rows=120000] [queries=2000] [rounds=6]
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/aslam_b200.h"

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

static bool read_all(int fd, void* p, size_t n)
{
    char* c = (char*)p;
    while (n) { ssize_t r = read(fd, c, n); if (r <= 0) return false; c += r; n -= (size_t)r; }
    return true;
}
static bool write_all(int fd, const void* p, size_t n)
{
    const char* c = (const char*)p;
    while (n) { ssize_t r = write(fd, c, n); if (r <= 0) return false; c += r; n -= (size_t)r; }
    return true;
}
#define MUST(call) do { if ((call) != ASLAM_OK) { fprintf(stderr, "rank %d: %s -> %s\n", rank, #call, aslam_last_error()); return 3; } } while (0)

static int child(int rank, int nranks, int rows, int nq, int rounds, int to_parent, int from_parent)
{
    const int kf_rows = 2000, nkf = (rows + kf_rows - 1) / kf_rows;
    std::vector<uint8_t> db((size_t)rows * 32), q((size_t)nq * 32);
    uint32_t s = 12345u;
    for (auto& b : db) b = (uint8_t)lcg(s);
    for (int i = 0; i < nq; ++i) {                              // queries = database rows with a few flipped bits
        const uint32_t src = lcg(s) % (uint32_t)rows;
        memcpy(&q[(size_t)i * 32], &db[(size_t)src * 32], 32);
        for (int f = 0; f < (int)(lcg(s) % 30); ++f) { const uint32_t bit = lcg(s) % 256; q[(size_t)i * 32 + bit / 8] ^= (uint8_t)(1u << (bit % 8)); }
        if (i % 16 == 15) memcpy(&q[(size_t)i * 32], &q[(size_t)(i - 1) * 32], 32);       // duplicates: ties
    }
    aslam_ctx* ctx = nullptr;
    MUST(aslam_ctx_create(rank, nullptr, &ctx));
    const int lo = nkf * rank / nranks, hi = nkf * (rank + 1) / nranks;      // this rank's keyframes
    aslam_loopdb* shard = nullptr;
    MUST(aslam_loopdb_create(ctx, 1000, nkf, &shard));          // small first allocation: the store grows
    for (int k = lo; k < hi; ++k) {
        const int r0 = k * kf_rows, n = (r0 + kf_rows <= rows ? kf_rows : rows - r0);
        MUST(aslam_loopdb_add(shard, &db[(size_t)r0 * 32], n, (uint64_t)k));
    }
    MUST(aslam_loopdb_set_row_base(shard, (int64_t)lo * kf_rows));
    MUST(aslam_loopdb_set_pipelining(shard, 1));
    // handles: export, send to the parent, receive everybody's
    std::vector<uint8_t> mine(aslam_loopdb_shard_handle_bytes()), all(mine.size() * (size_t)nranks);
    MUST(aslam_loopdb_shard_export(shard, nranks, nq, mine.data()));
    if (!write_all(to_parent, mine.data(), mine.size()) || !read_all(from_parent, all.data(), all.size())) return 4;
    MUST(aslam_loopdb_shard_init(shard, rank, nranks, all.data()));
    char go = 0;
    if (!write_all(to_parent, "r", 1) || !read_all(from_parent, &go, 1)) return 4;         // every rank has opened every buffer

    aslam_loopdb* whole = nullptr;
    std::vector<uint64_t> ref((size_t)nq * 2, 0), got((size_t)nq * 2, 0);
    if (rank == 0) {
        MUST(aslam_loopdb_create(ctx, rows, nkf, &whole));
        for (int k = 0; k < nkf; ++k) {
            const int r0 = k * kf_rows, n = (r0 + kf_rows <= rows ? kf_rows : rows - r0);
            MUST(aslam_loopdb_add(whole, &db[(size_t)r0 * 32], n, (uint64_t)k));
        }
        MUST(aslam_loopdb_query_global(whole, q.data(), nq, ref.data()));
    }
    // This program does not link the CUDA runtime, so the query block and the result live in pinned host memory from
    // aslam_host_alloc, which the device addresses directly (unified virtual addressing): the kernel reads the queries
    // and writes the keys through PCIe.  Slower than device buffers, and exactly as correct.
    void *d_q_host = nullptr, *d_out_host = nullptr;
    MUST(aslam_host_alloc((size_t)nq * 32, &d_q_host));
    MUST(aslam_host_alloc((size_t)nq * 16, &d_out_host));
    memcpy(d_q_host, q.data(), q.size());
    int bad = 0;
    for (int r = 0; r < rounds; ++r) {
        memset(d_out_host, 0, (size_t)nq * 16);
        MUST(aslam_loopdb_query_global_sharded_device(shard, (const uint8_t*)d_q_host, nq, (uint64_t*)d_out_host));
        if (r % 2 == 1) {                                       // every other round back to back (pipelined), then compare
            MUST(aslam_ctx_sync(ctx));
            memcpy(got.data(), d_out_host, got.size() * 8);
            if (rank == 0 && memcmp(got.data(), ref.data(), got.size() * 8) != 0) ++bad;
            // all ranks must hold the same keys: send a checksum to the parent
            uint64_t sum = 1469598103934665603ull;
            for (uint64_t v : got) { sum ^= v; sum *= 1099511628211ull; }
            if (!write_all(to_parent, &sum, 8)) return 4;
        }
    }
    MUST(aslam_ctx_sync(ctx));
    uint32_t status = 0;
    MUST(aslam_loopdb_exchange_status(shard, &status, nullptr, 0, 0));
    aslam_host_free(d_q_host); aslam_host_free(d_out_host);
    aslam_loopdb_destroy(whole);
    aslam_loopdb_destroy(shard);
    aslam_ctx_destroy(ctx);
    if (status) { fprintf(stderr, "rank %d: exchange timed out\n", rank); return 5; }
    if (bad) { fprintf(stderr, "rank 0: %d rounds differ from the unsharded database\n", bad); return 6; }
    return 0;
}

int main(int argc, char** argv)
{
    const int nranks = argc > 1 ? atoi(argv[1]) : 2, rows = argc > 2 ? atoi(argv[2]) : 120000;
    const int nq = argc > 3 ? atoi(argv[3]) : 2000, rounds = argc > 4 ? atoi(argv[4]) : 6;
    if (nranks < 1 || nranks > 16) { fprintf(stderr, "1..16 ranks\n"); return 2; }
    std::vector<int> up((size_t)nranks * 2), down((size_t)nranks * 2);
    std::vector<pid_t> pid((size_t)nranks);
    for (int r = 0; r < nranks; ++r) {
        if (pipe(&up[(size_t)r * 2]) || pipe(&down[(size_t)r * 2])) return 2;
        pid[(size_t)r] = fork();                                // before any CUDA call in this process
        if (pid[(size_t)r] == 0) {
            close(up[(size_t)r * 2]); close(down[(size_t)r * 2 + 1]);
            _exit(child(r, nranks, rows, nq, rounds, up[(size_t)r * 2 + 1], down[(size_t)r * 2]));
        }
        close(up[(size_t)r * 2 + 1]); close(down[(size_t)r * 2]);
    }
    const size_t hb = 64;                                       // aslam_loopdb_shard_handle_bytes(); the parent stays CUDA-free
    std::vector<uint8_t> all(hb * (size_t)nranks);
    bool ok = true;
    for (int r = 0; r < nranks; ++r) ok = ok && read_all(up[(size_t)r * 2], &all[hb * (size_t)r], hb);
    for (int r = 0; r < nranks; ++r) ok = ok && write_all(down[(size_t)r * 2 + 1], all.data(), all.size());
    char c = 0;
    for (int r = 0; r < nranks; ++r) ok = ok && read_all(up[(size_t)r * 2], &c, 1);
    for (int r = 0; r < nranks; ++r) ok = ok && write_all(down[(size_t)r * 2 + 1], "g", 1);
    int mismatched = 0;
    for (int k = 0; ok && k < rounds / 2; ++k) {
        uint64_t first = 0;
        for (int r = 0; r < nranks; ++r) {
            uint64_t sum = 0;
            ok = ok && read_all(up[(size_t)r * 2], &sum, 8);
            if (r == 0) first = sum; else if (sum != first) ++mismatched;
        }
    }
    int failed = 0;
    for (int r = 0; r < nranks; ++r) {
        int st = 0;
        waitpid(pid[(size_t)r], &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) { fprintf(stderr, "rank %d exited with %d\n", r, WIFEXITED(st) ? WEXITSTATUS(st) : -1); ++failed; }
    }
    if (!ok || failed || mismatched) { fprintf(stderr, "shard_ipc_selftest FAILED (pipes ok %d, failed ranks %d, rounds with differing ranks %d)\n", (int)ok, failed, mismatched); return 1; }
    printf("shard_ipc_selftest ok: %d processes, %d rows, %d queries, %d rounds; every rank equals the unsharded database\n", nranks, rows, nq, rounds);
    return 0;
}
