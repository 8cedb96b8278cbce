// test_multi_gpu.cpp — one rank of a multi-GPU run written against the C++ host mirror (include/mdim/view.hpp) and the
// multi-GPU surface of the C ABI (mdim_comm_init / mdim_peer_table / mdim_allgather / mdim_allreduce): no Python, no
// torch.distributed — what the north star's "Rust host -> extern C" caller would do, in the one typed host language
// this image can compile.
//   ./test_multi_gpu <rank> <world> <libmdim_b200.so> <rendezvous file>
#include <dlfcn.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>

#include "../../include/mdim/view.hpp"

using namespace mdim;
using U2 = std::tuple<usize, usize>;

static int failures = 0, checks = 0;
#define CHECK(cond) do { ++checks; if (!(cond)) { ++failures; std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } } while (0)

int main(int argc, char** argv) {
    if (argc < 5) { std::printf("usage: test_multi_gpu <rank> <world> <library> <rendezvous file>\n"); return 2; }
    const int rank = std::atoi(argv[1]), world = std::atoi(argv[2]);
    void* lib = dlopen(argv[3], RTLD_NOW);
    if (!lib) { std::printf("dlopen: %s\n", dlerror()); return 2; }
    const Api api = Api::load([&](const char* name) { void* p = dlsym(lib, name); if (!p) { std::printf("missing symbol %s\n", name); std::exit(2); } return p; });
    mdim_ctx* ctx = nullptr;
    if (api.init(rank, &ctx) != MDIM_OK) { std::printf("mdim_init(%d) failed\n", rank); return 2; }

    // rendezvous: rank 0 makes the communicator id and publishes its 128 bytes in a file; the others wait for it
    uint8_t id[MDIM_COMM_ID_BYTES];
    const std::string path = argv[4];
    if (rank == 0) {
        api.check(ctx, api.comm_unique_id(id));
        { std::ofstream f(path + ".tmp", std::ios::binary); f.write((const char*)id, sizeof id); }
        std::rename((path + ".tmp").c_str(), path.c_str());
    } else {
        for (int tries = 0; tries < 60000; ++tries) { std::ifstream f(path, std::ios::binary); if (f && f.read((char*)id, sizeof id)) break; usleep(5000); }
    }
    api.check(ctx, api.comm_init(ctx, rank, world, id));

    // a (M, N) f32 matrix sharded by rows: this rank owns rows [rank * M / world, ...)
    const uint64_t M = 128 * (uint64_t)world, N = 320, rows = M / (uint64_t)world, block = rows * N;
    auto value = [&](uint64_t y, uint64_t x) { return (float)(y * N + x) * 0.5f; };
    std::vector<float> mine(block);
    for (uint64_t y = 0; y < rows; ++y) for (uint64_t x = 0; x < N; ++x) mine[y * N + x] = value((uint64_t)rank * rows + y, x);
    DeviceArray<U2, float> local(api, ctx, std::make_tuple(rows, N), mine);
    api.check(ctx, api.barrier(ctx));  // every block is filled before anyone reads a peer

    // (1) transpose of the row-sharded Array, this rank's block of the transposed rows: every tile is read from the owning GPU
    void* peers_raw[MDIM_MAX_PEERS];
    api.check(ctx, api.peer_table(ctx, local.device_ptr(), block * 4, peers_raw));
    std::vector<const void*> peers(peers_raw, peers_raw + world);
    auto whole = DeviceArray<U2, float>::sharded(std::make_tuple(M, N), peers, block);
    auto mine_t = collect_device(whole.transpose<Unit, usize, usize, Unit>().iso<U2>().shard(rank, world), api, ctx).to_raw();
    const uint64_t xq = N / (uint64_t)world, xr = N % (uint64_t)world, xlo = (uint64_t)rank * xq + std::min<uint64_t>((uint64_t)rank, xr), xn = xq + ((uint64_t)rank < xr ? 1 : 0);
    bool same = mine_t.size() == xn * M;
    for (uint64_t x = 0; x < xn && same; ++x) for (uint64_t y = 0; y < M; ++y) same = same && mine_t[x * M + y] == value(y, xlo + x);
    CHECK(same);

    // (2) the north star's route: all-gather the source, then the local block transpose
    DeviceArray<U2, float> full(api, ctx, std::make_tuple(M, N));
    api.check(ctx, api.allgather(ctx, local.device_ptr(), full.device_ptr(), block * 4));
    auto mine_t2 = collect_device(full.transpose<Unit, usize, usize, Unit>().iso<U2>().shard(rank, world), api, ctx).to_raw();
    CHECK(mine_t2 == mine_t);

    // (3) fold over the SHARDED axis: per-rank partial fold (sequential over the local rows) + all-reduce
    auto part = collect_device(local.transpose<Unit, usize, usize, Unit>().iso<U2>().rows<usize, usize>().fold<Add>(0.0f), api, ctx);
    api.check(ctx, api.allreduce(ctx, part.device_ptr(), N, MDIM_F32, MDIM_ADD));
    api.check(ctx, api.sync(ctx));
    auto sums = part.to_raw();
    same = sums.size() == N;
    for (uint64_t x = 0; x < N && same; ++x) {
        double want = 0; for (uint64_t y = 0; y < M; ++y) want += (double)value(y, x);
        same = same && std::abs((double)sums[x] - want) <= 1e-6 * std::abs(want);
    }
    CHECK(same);

    // (4) the same fold as ONE fused kernel per GPU, both routes: the bit-exact chain through the ranks (k_fold_ring) equals the
    //     sequential f32 sum over ALL rows in index order; the blocked route (k_fold_xchg) equals the per-rank sequential partial sums
    //     (rank 0 from the initial value, the others from -0.0) combined in rank order — both restated here, bit for bit
    {
        DeviceArray<usize, float> ring(api, ctx, N), blocked(api, ctx, N);
        mdim_scalar init; std::memset(&init, 0, sizeof init); init.f32 = 0.5f;
        api.check(ctx, api.fold_sharded_axis(ctx, local.device_ptr(), M / (uint64_t)world, N, MDIM_F32, MDIM_ADD, init, ring.device_ptr()));
        api.check(ctx, api.fold_sharded_axis_blocked(ctx, local.device_ptr(), M / (uint64_t)world, N, MDIM_F32, MDIM_ADD, init, blocked.device_ptr()));
        api.check(ctx, api.fold_sharded_axis_status(ctx));
        api.check(ctx, api.sync(ctx));
        auto got_ring = ring.to_raw(), got_blocked = blocked.to_raw();
        bool ring_ok = got_ring.size() == N, blocked_ok = got_blocked.size() == N;
        const uint64_t rows_per_rank = M / (uint64_t)world;
        for (uint64_t x = 0; x < N && (ring_ok || blocked_ok); ++x) {
            volatile float chain = 0.5f, total = 0.0f;
            for (uint64_t y = 0; y < M; ++y) chain = chain + value(y, x);
            for (int r = 0; r < world; ++r) {
                volatile float part = r == 0 ? 0.5f : -0.0f;
                for (uint64_t y = (uint64_t)r * rows_per_rank; y < (uint64_t)(r + 1) * rows_per_rank; ++y) part = part + value(y, x);
                total = r == 0 ? part : total + part;
            }
            float c = chain, t = total;
            ring_ok = ring_ok && std::memcmp(&got_ring[x], &c, 4) == 0;
            blocked_ok = blocked_ok && std::memcmp(&got_blocked[x], &t, 4) == 0;
        }
        CHECK(ring_ok);
        CHECK(blocked_ok);
    }

    api.check(ctx, api.peer_table_close(ctx));
    api.check(ctx, api.comm_destroy(ctx));
    std::printf("rank %d: %d checks, %d failures\n", rank, checks, failures);
    api.shutdown(ctx);
    return failures ? 1 : 0;
}
