// Multi-GPU arg-min through the C ABI alone (no Python, no torch): one thread per GPU, each solves its shard of a batch of
// candidate corridors with frx_batch_optimize and calls frx_batch_argmin_allgather on a communicator from ncclCommInitAll.
// Every rank must report the same (cost, global index), equal to the minimum over the per-rank arg-mins.
// usage: argmin_nccl [n_gpus]   (default: all visible devices)
#include <cuda_runtime.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "frx.h"

static std::vector<double> box(double cx, double cy, double cz, double hx, double hy, double hz) {
    const double n[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    const double h[6] = {hx, hx, hy, hy, hz, hz};
    std::vector<double> f;
    for (int k = 0; k < 6; k++) {
        for (int d = 0; d < 3; d++) f.push_back(n[k][d]);
        f.push_back(cx + n[k][0] * h[k]), f.push_back(cy + n[k][1] * h[k]), f.push_back(cz + n[k][2] * h[k]);
    }
    return f;
}

int main(int argc, char** argv) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return 2;
    if (argc > 1) ndev = std::min(ndev, atoi(argv[1]));
    const int per_rank = 3;
    std::vector<ncclComm_t> comms(ndev);
    std::vector<int> devs(ndev);
    for (int i = 0; i < ndev; i++) devs[i] = i;
    if (ncclCommInitAll(comms.data(), ndev, devs.data()) != ncclSuccess) return 3;
    std::vector<long long> best_idx(ndev, -1);
    std::vector<double> best_cost(ndev, 0.0), local_cost(ndev, 0.0);
    std::vector<int> status(ndev, 0);
    std::vector<std::thread> pool;
    for (int r = 0; r < ndev; r++)
        pool.emplace_back([&, r]() {
            cudaSetDevice(r);
            // candidates differ by the lateral offset of the middle boxes: different corridors, different costs
            std::vector<std::vector<double>> faces(per_rank);
            std::vector<std::vector<int>> off(per_rank);
            std::vector<frx_problem> probs(per_rank);
            for (int c = 0; c < per_rank; c++) {
                const int g = r * per_rank + c;
                off[c] = {0};
                for (int i = 0; i < 4; i++) {
                    const auto b = box(3.0 * i + 1.5, 0.15 * g * (i == 1 || i == 2), 1.5, 3.0, 2.5, 1.4);
                    faces[c].insert(faces[c].end(), b.begin(), b.end());
                    off[c].push_back(off[c].back() + 6);
                }
                frx_problem& p = probs[c];
                p = frx_problem{};
                const double ini[9] = {0, 0, 1.5, 0, 0, 0, 0, 0, 0}, fin[9] = {12.0, 0.0, 1.5, 0, 0, 0, 0, 0, 0};
                for (int k = 0; k < 9; k++) p.ini[k] = ini[k], p.fin[k] = fin[k];
                p.n_poly = 4, p.face_off = off[c].data(), p.faces = faces[c].data();
                p.params = frx_params{1000.0, 0.0, INFINITY, 16, 1, 0.5, 0.15, 0.08, 14.0, 5.0, 12.0, 3.8, 9.81, {1.0e7, 1.0e4, 1.0e4, 1.0e4}};
            }
            frx_batch* b = nullptr;
            int s = frx_batch_create(probs.data(), per_rank, r, &b);
            if (s == FRX_OK) s = frx_batch_optimize(b, 1.0e-5, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
            int li = -1;
            if (s == FRX_OK) s = frx_batch_argmin(b, &li, &local_cost[r]);
            if (s == FRX_OK) s = frx_batch_argmin_allgather(b, comms[r], ndev, (long long)r * per_rank, &best_idx[r], &best_cost[r]);
            status[r] = s;
            frx_batch_destroy(b);
        });
    for (auto& t : pool) t.join();
    for (auto c : comms) ncclCommDestroy(c);
    double expect = INFINITY;
    bool ok = true;
    for (int r = 0; r < ndev; r++) ok = ok && status[r] == FRX_OK, expect = std::min(expect, local_cost[r]);
    for (int r = 0; r < ndev; r++) ok = ok && best_idx[r] == best_idx[0] && best_cost[r] == best_cost[0];
    ok = ok && best_cost[0] == expect && best_idx[0] >= 0 && best_idx[0] < (long long)ndev * per_rank;
    std::printf("ranks %d: arg-min cost %.9f at global index %lld (%s)\n", ndev, best_cost[0], best_idx[0], ok ? "all ranks agree" : "MISMATCH");
    return ok ? 0 : 1;
}
