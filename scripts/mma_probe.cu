// mma_probe.cu — standalone check + timing of knn2_mma_kernel (uz_knn2_mma.cuh) against a scalar CPU kNN-2 and against
// knn2_kernel (the POPC form) on the same descriptors.  Build: scripts/build_mma_probe.sh; run under gpurun.
//   mma_probe check            small shapes, bit-exact against the CPU loop, for the descriptor variants given by --variant
//   mma_probe time P N         P pairs of N x N over a keyframe pool, both kernels, ms per launch
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../uzliti_slam_b200/csrc/uz_knn2_mma.cuh"
#include "../uzliti_slam_b200/csrc/uz_knn2_mma2.cuh"
#include "../uzliti_slam_b200/csrc/uz_knn2_mmak.cuh"
#include "../uzliti_slam_b200/csrc/uz_knn2_mmaf.cuh"

using namespace uz;

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(2); } } while (0)

static void cpu_knn2(const uint32_t* q, int nq, const uint32_t* t, int nt, std::vector<uint2>& out) {
    out.assign(nq, make_uint2(kNoKey, kNoKey));
    for (int i = 0; i < nq; ++i) {
        uint32_t m1 = kNoKey, m2 = kNoKey;
        for (int j = 0; j < nt; ++j) {
            int d = 0;
            for (int w = 0; w < 8; ++w) d += __builtin_popcount(q[i * 8 + w] ^ t[j * 8 + w]);
            const uint32_t k = ((uint32_t)d << 16) | (uint32_t)j;
            if (k < m1) { m2 = m1; m1 = k; } else if (k < m2) m2 = k;
        }
        out[i] = make_uint2(m1, m2);
    }
}

static const uint8_t* zero_page() {
    static uint8_t* z = nullptr;
    if (!z) { CK(cudaMalloc(&z, kF4ZeroPageBytes)); CK(cudaMemset(z, 0, kF4ZeroPageBytes)); }
    return z;
}
struct Cam { uint32_t* raw; uint32_t* csa; uint8_t* e8; uint8_t* e4; int n; std::vector<uint32_t> h; };

static Cam make_cam(int n, std::mt19937& rng, int mode) {
    Cam c; c.n = n; c.h.resize((size_t)std::max(n, 1) * 8);
    for (auto& w : c.h) w = rng();
    if (mode == 1) for (int i = 0; i < n; ++i) for (int w = 1; w < 8; ++w) c.h[i * 8 + w] = 0;          // tie-heavy
    if (mode == 2) for (int i = 1; i < n; i += 3) memcpy(&c.h[i * 8], &c.h[(i - 1) * 8], 32);          // duplicates
    CK(cudaMalloc(&c.raw, (size_t)std::max(n, 1) * 32));
    CK(cudaMalloc(&c.csa, (size_t)std::max(n, 1) * 32));
    CK(cudaMalloc(&c.e8, e8_bytes(std::max(n, 1))));
    CK(cudaMemset(c.e8, 0x7F, e8_bytes(std::max(n, 1))));     // garbage in the padding rows on purpose
    CK(cudaMalloc(&c.e4, e4_bytes(std::max(n, 1))));
    CK(cudaMemset(c.e4, 0x7F, e4_bytes(std::max(n, 1))));
    if (n) {
        CK(cudaMemcpy(c.raw, c.h.data(), (size_t)n * 32, cudaMemcpyHostToDevice));
        pack_descriptors_kernel<<<(n + 255) / 256, 256>>>((const uint8_t*)c.raw, n, 32, c.raw, c.csa, 1);
        expand_e8_kernel<<<(n * 16 + 255) / 256, 256>>>(c.raw, n, c.e8);
        expand_e4_kernel<<<(((n + 7) & ~7) * 8 + 255) / 256, 256>>>(c.raw, n, c.e4);
        CK(cudaGetLastError());
    }
    return c;
}

static int g_sms = 148;

static float run_mma(const std::vector<MmaTask>& tasks, size_t key_rows, MmaDesc dsc, std::vector<uint2>* out, int reps, int variant = 0) {
    std::vector<int2> items;
    for (size_t t = 0; t < tasks.size(); ++t)
        for (int q0 = 0; q0 < tasks[t].nq; q0 += kMmaItemRows) items.push_back(make_int2((int)t, q0));
    MmaTask* d_tasks; int2* d_items; uint2* d_keys;
    CK(cudaMalloc(&d_tasks, std::max<size_t>(tasks.size(), 1) * sizeof(MmaTask)));
    CK(cudaMalloc(&d_items, std::max<size_t>(items.size(), 1) * sizeof(int2)));
    CK(cudaMalloc(&d_keys, std::max<size_t>(key_rows, 1) * sizeof(uint2)));
    CK(cudaMemcpy(d_tasks, tasks.data(), tasks.size() * sizeof(MmaTask), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_items, items.data(), items.size() * sizeof(int2), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_keys, 0xEE, std::max<size_t>(key_rows, 1) * sizeof(uint2)));
    CK(cudaFuncSetAttribute(knn2_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmaSmemBytes));
    CK(cudaFuncSetAttribute(knn2_mmak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmakSmemBytes));
    CK(cudaFuncSetAttribute(knn2_mmaf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kF4SmemBytes));
    const int grid = (int)std::min<size_t>(items.size(), (size_t)g_sms);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        if (grid > 0 && variant == 5) knn2_mmaf_kernel<false><<<grid, kF4Threads, kF4SmemBytes>>>(d_tasks, d_items, (int)items.size(), d_keys, dsc, nullptr, nullptr, zero_page());
        else if (grid > 0 && variant == 2) knn2_mmak_kernel<<<grid, kMmaThreads, kMmakSmemBytes>>>(d_tasks, d_items, (int)items.size(), d_keys, dsc, nullptr, nullptr);
        else if (grid > 0) knn2_mma_kernel<<<grid, kMmaThreads, kMmaSmemBytes>>>(d_tasks, d_items, (int)items.size(), d_keys, dsc, nullptr, nullptr);
        cudaEventRecord(e1);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("knn2_mma_kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 || reps == 1) best = std::min(best, ms);
    }
    if (out) { out->resize(key_rows); CK(cudaMemcpy(out->data(), d_keys, key_rows * sizeof(uint2), cudaMemcpyDeviceToHost)); }
#ifdef UZ_F4_TRACE
    if (variant == 5) {
        static long long h[2][128][8];
        CK(cudaMemcpyFromSymbol(h, g_f4_trace, sizeof(h)));
        const long long z = h[0][0][0];
        printf("  acc | issuer: tile loop top, tile seen, free seen, committed | epilogue: full seen, read + released, swept   (clocks since the first issue)\n");
        for (int a = 40; a < 72; ++a)
            printf("  %3d | %7lld %7lld %7lld %7lld | %7lld %7lld %7lld\n", a, h[0][a][2] - z, h[0][a][3] - z, h[0][a][0] - z,
                   h[0][a][1] - z, h[1][a][0] - z, h[1][a][1] - z, h[1][a][2] - z);
    }
#endif
#ifdef UZ_MMA_PROF
    {
        static long long h[256][4];
        CK(cudaMemcpyFromSymbol(h, g_mma_prof, sizeof(h)));
        double s[4] = {0, 0, 0, 0};
        for (int b = 0; b < grid; ++b) for (int k = 0; k < 4; ++k) s[k] += (double)h[b][k] / grid;
        printf("  MMA issuer, mean clocks per CTA: loop %.0f, waiting for train tiles %.0f (%.1f %%), query tiles %.0f (%.1f %%), free accumulators %.0f (%.1f %%)\n",
               s[3], s[0], 100 * s[0] / s[3], s[1], 100 * s[1] / s[3], s[2], 100 * s[2] / s[3]);
    }
#endif
    cudaFree(d_tasks); cudaFree(d_items); cudaFree(d_keys);
    return best;
}

static float run_mma2(const std::vector<MmaTask>& tasks, size_t key_rows, MmaDesc dsc, std::vector<uint2>* out, int reps) {
    std::vector<int2> items;
    for (size_t t = 0; t < tasks.size(); ++t)
        for (int q0 = 0; q0 < tasks[t].nq; q0 += kMma2ItemRows) items.push_back(make_int2((int)t, q0));
    MmaTask* d_tasks; int2* d_items; uint2* d_keys;
    CK(cudaMalloc(&d_tasks, std::max<size_t>(tasks.size(), 1) * sizeof(MmaTask)));
    CK(cudaMalloc(&d_items, std::max<size_t>(items.size(), 1) * sizeof(int2)));
    CK(cudaMalloc(&d_keys, std::max<size_t>(key_rows, 1) * sizeof(uint2)));
    CK(cudaMemcpy(d_tasks, tasks.data(), tasks.size() * sizeof(MmaTask), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_items, items.data(), items.size() * sizeof(int2), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_keys, 0xEE, std::max<size_t>(key_rows, 1) * sizeof(uint2)));
    CK(cudaFuncSetAttribute(knn2_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMma2SmemBytes));
    const int clusters = (int)std::min<size_t>(items.size(), (size_t)g_sms / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(kMmaThreads); cfg.dynamicSmemBytes = kMma2SmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        if (clusters > 0) CK(cudaLaunchKernelEx(&cfg, knn2_mma2_kernel, (const MmaTask*)d_tasks, (const int2*)d_items, (int)items.size(), d_keys, dsc));
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("knn2_mma2_kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 || reps == 1) best = std::min(best, ms);
    }
    if (out) { out->resize(key_rows); CK(cudaMemcpy(out->data(), d_keys, key_rows * sizeof(uint2), cudaMemcpyDeviceToHost)); }
#ifdef UZ_MMA_PROF
    {
        static long long h[256][4];
        CK(cudaMemcpyFromSymbol(h, g_mma_prof, sizeof(h)));
        double s[4] = {0, 0, 0, 0};
        for (int b = 0; b < clusters; ++b) for (int k = 0; k < 4; ++k) s[k] += (double)h[b][k] / clusters;
        printf("  pair kernel issuer, mean clocks per cluster: loop %.0f, waiting for train tiles %.0f (%.1f %%), query tiles %.0f (%.1f %%), free accumulators %.0f (%.1f %%)\n",
               s[3], s[0], 100 * s[0] / s[3], s[1], 100 * s[1] / s[3], s[2], 100 * s[2] / s[3]);
    }
#endif
    cudaFree(d_tasks); cudaFree(d_items); cudaFree(d_keys);
    return best;
}

static float run_popc(const std::vector<MatchTask>& tasks, size_t key_rows, std::vector<uint2>* out, int reps) {
    std::vector<int2> tiles;
    for (size_t t = 0; t < tasks.size(); ++t)
        for (int q0 = 0; q0 < tasks[t].nq; q0 += 512) tiles.push_back(make_int2((int)t, q0));
    MatchTask* d_tasks; int2* d_tiles; uint2* d_keys;
    CK(cudaMalloc(&d_tasks, std::max<size_t>(tasks.size(), 1) * sizeof(MatchTask)));
    CK(cudaMalloc(&d_tiles, std::max<size_t>(tiles.size(), 1) * sizeof(int2)));
    CK(cudaMalloc(&d_keys, std::max<size_t>(key_rows, 1) * sizeof(uint2)));
    CK(cudaMemcpy(d_tasks, tasks.data(), tasks.size() * sizeof(MatchTask), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        if (!tiles.empty()) knn2_kernel<256, 2, true, true><<<(unsigned)tiles.size(), 256, knn_smem_bytes(256, 2)>>>(d_tasks, d_tiles, d_keys, nullptr, nullptr);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 || reps == 1) best = std::min(best, ms);
    }
    if (out) { out->resize(key_rows); CK(cudaMemcpy(out->data(), d_keys, key_rows * sizeof(uint2), cudaMemcpyDeviceToHost)); }
    cudaFree(d_tasks); cudaFree(d_tiles); cudaFree(d_keys);
    return best;
}

static float run_mma2(const std::vector<MmaTask>& tasks, size_t key_rows, MmaDesc dsc, std::vector<uint2>* out, int reps);
static int check(MmaDesc dsc, const char* name) {
    std::mt19937 rng(1234);
    const int shapes[][3] = {{128, 256, 0}, {256, 256, 0}, {500, 500, 0}, {1000, 1000, 0}, {1000, 1000, 1}, {777, 333, 2},
                             {1, 1, 0}, {5, 2, 0}, {300, 0, 0}, {129, 257, 0}, {4096, 4096, 0}, {260, 1031, 1}, {1000, 31, 0}, {9, 700, 2}};
    int bad_total = 0;
    std::vector<Cam> cams;
    std::vector<MmaTask> tasks;
    std::vector<std::vector<uint2>> want;
    size_t key_rows = 0;
    for (auto& s : shapes) {
        Cam q = make_cam(s[0], rng, s[2]), t = make_cam(s[1], rng, s[2]);
        MmaTask tk; memset(&tk, 0, sizeof(tk)); tk.q_desc = (const uint32_t*)q.e8; tk.t_desc = (const uint32_t*)t.e8; tk.nq = s[0]; tk.nt = s[1]; tk.key_off = (uint32_t)key_rows; tk.pair = (int)tasks.size();
        key_rows += (size_t)s[0];
        tasks.push_back(tk);
        want.emplace_back();
        cpu_knn2(q.h.data(), s[0], t.h.data(), s[1], want.back());
        cams.push_back(std::move(q)); cams.push_back(std::move(t));
    }
    CK(cudaDeviceSynchronize());
    std::vector<uint2> got;
    const int variant = getenv("PROBE_VARIANT") ? atoi(getenv("PROBE_VARIANT")) : 0;
    if (variant == 5) for (size_t k = 0; k < tasks.size(); ++k) { tasks[k].q_desc = (const uint32_t*)cams[2 * k].e4; tasks[k].t_desc = (const uint32_t*)cams[2 * k + 1].e4; }
    if (variant == 3) run_mma2(tasks, key_rows, dsc, &got, 1); else run_mma(tasks, key_rows, dsc, &got, 1, variant);
    for (size_t k = 0; k < tasks.size(); ++k) {
        int bad = 0;
        for (int i = 0; i < tasks[k].nq; ++i) {
            const uint2 g = got[tasks[k].key_off + i], w = want[k][i];
            if (g.x != w.x || g.y != w.y) {
                if (bad < 3) printf("  [%s] shape %dx%d row %d: got (%08x,%08x) want (%08x,%08x)\n", name, tasks[k].nq, tasks[k].nt, i, g.x, g.y, w.x, w.y);
                ++bad;
            }
        }
        printf("[%s] %4d x %4d mode %d: %s (%d bad rows)\n", name, tasks[k].nq, tasks[k].nt, shapes[k][2], bad ? "MISMATCH" : "ok", bad);
        bad_total += bad;
    }
    for (auto& c : cams) { cudaFree(c.raw); cudaFree(c.csa); cudaFree(c.e8); cudaFree(c.e4); }
    return bad_total;
}

int main(int argc, char** argv) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, cc %d.%d\n", prop.name, g_sms, prop.major, prop.minor);
    MmaDesc dsc = uz_knn2_mma_desc();
    const char* mode = argc > 1 ? argv[1] : "check";
    for (int a = 2; a < argc; ++a) if (!strcmp(argv[a], "--swap")) std::swap(dsc.lbo16, dsc.sbo16);
    if (!strcmp(mode, "check")) {
        const int bad = check(dsc, "mma");
        printf("CHECK %s\n", bad ? "FAILED" : "PASSED");
        return bad ? 1 : 0;
    }
    // time P N [pool]
    const int P = argc > 2 ? atoi(argv[2]) : 5000, N = argc > 3 ? atoi(argv[3]) : 1000;
    const int pool = argc > 4 && argv[4][0] != '-' ? atoi(argv[4]) : std::max(2, P / 10);
    std::mt19937 rng(99);
    std::vector<Cam> cams;
    for (int i = 0; i < pool; ++i) cams.push_back(make_cam(N, rng, 0));
    std::vector<MmaTask> mt; std::vector<MatchTask> pt;
    std::vector<int> tq_of, tf_of;
    size_t key_rows = 0;
    for (int p = 0; p < P; ++p) {
        const int f = (p / 20) % pool, t = (int)(rng() % pool);        // 20 candidates per from-keyframe, as in C4
        MmaTask a; memset(&a, 0, sizeof(a)); a.q_desc = (const uint32_t*)cams[t].e8; a.t_desc = (const uint32_t*)cams[f].e8; a.nq = N; a.nt = N; a.key_off = (uint32_t)key_rows; a.pair = p;
        MatchTask b; memset(&b, 0, sizeof(b));
        b.q_desc = cams[t].csa; b.t_desc = cams[f].csa; b.nq = N; b.nt = N; b.key_off = (uint32_t)key_rows; b.pair = p; b.rev_key_off = kNoRev;
        key_rows += N;
        mt.push_back(a); pt.push_back(b); tq_of.push_back(t); tf_of.push_back(f);
    }
    CK(cudaDeviceSynchronize());
    std::vector<uint2> k_mma, k_popc;
    const float ms_popc = run_popc(pt, key_rows, &k_popc, 4);
    const float ms_mma = run_mma(mt, key_rows, dsc, &k_mma, 4);
    std::vector<uint2> kk;
    const float ms_k = run_mma(mt, key_rows, dsc, &kk, 4, 2);
    size_t diffk = 0;
    for (size_t i = 0; i < key_rows; ++i) diffk += (kk[i].x != k_popc[i].x || kk[i].y != k_popc[i].y);
    printf("keys from the MMA: %.3f ms, rows differing %zu\n", ms_k, diffk);
    std::vector<MmaTask> ft = mt;
    for (size_t k = 0; k < ft.size(); ++k) { ft[k].q_desc = (const uint32_t*)cams[tq_of[k]].e4; ft[k].t_desc = (const uint32_t*)cams[tf_of[k]].e4; }
    std::vector<uint2> kf;
    const float ms_f = run_mma(ft, key_rows, dsc, &kf, 4, 5);
    size_t difff = 0;
    for (size_t i = 0; i < key_rows; ++i) difff += (kf[i].x != k_popc[i].x || kf[i].y != k_popc[i].y);
    printf("4-bit operands (kind::mxf4): %.3f ms, rows differing %zu\n", ms_f, difff);
    std::vector<uint2> k2a, k2b;
    const float ms_2a = run_mma2(mt, key_rows, dsc, &k2a, 4);
    k2b = k2a;
    size_t diff = 0, diff2 = 0;
    for (size_t i = 0; i < key_rows; ++i) diff += (k_mma[i].x != k_popc[i].x || k_mma[i].y != k_popc[i].y);
    for (size_t i = 0; i < key_rows; ++i) diff2 += (k2a[i].x != k_popc[i].x || k2a[i].y != k_popc[i].y) + (k2b[i].x != k_popc[i].x || k2b[i].y != k_popc[i].y);
    printf("pair kernel: %.3f ms, rows differing %zu\n", ms_2a, diff2 / 2);
    const double cmp = (double)P * N * N;
    printf("TIME pairs=%d N=%d pool=%d: popc %.3f ms (%.0f G cmp/s)  mma %.3f ms (%.0f G cmp/s)  speedup %.2fx  rows differing %zu of %zu\n",
           P, N, pool, ms_popc, cmp / ms_popc * 1e-6, ms_mma, cmp / ms_mma * 1e-6, ms_popc / ms_mma, diff, key_rows);
    return diff ? 1 : 0;
}
