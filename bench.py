#!/usr/bin/env python
"""bench.py — edge estimates/sec of the feature-edge path (Hamming kNN-2 + ratio + RANSAC rigid transform).

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[3], SURVEY.md §8d "C4"): a 10 000-keyframe map, 1000 ORB-256 features per
keyframe, K=20 loop-closure candidates per keyframe = 200 000 keyframe pairs, sharded over 8 GPUs.  One
"step" = one pass of the hot path over one rank's shard of 25 000 pairs (weak scaling: the shard size is
fixed, the map is replicated on every GPU, at N=8 the step is exactly C4).  Pairs are fully independent;
the only exchange is the all-gather of the 176-byte per-pair edge records over NCCL.

The JSON line carries:  value = whole-job edges/s with the keyframe store resident in HBM;  e2e = the same
batch through uz_estimate_edges_host() from pinned HOST buffers (H2D of every referenced keyframe + D2H of
the edge records inside the timed region);  roofline for the dominant kernel (knn2) against integer-pipe
peaks microbenchmarked in this same run;  cpu_baseline = the oracle port timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_KEYFRAMES = 10000
N_FEATURES = 1000
K_CAND = 20
PAIRS_PER_GPU = 25000          # 200 000 / 8
STREAM_SOLVE = int(os.environ.get("UZ_STREAM_SOLVE", "1"))   # persistent solve CTAs per SM beside the match kernel (0 = solve behind it)
METRIC = "edge estimates/sec (Hamming kNN-2 + ratio + RANSAC)"
UNIT = "edges/s"
WORKLOAD = ("C4 loop-closure screening: 10000-keyframe map x 1000 ORB-256 features, K=20 candidates/keyframe, "
            "25000 keyframe pairs per GPU (200000 pairs at 8 GPUs), 100 RANSAC hypotheses/pair, thr 0.1 m, break 0.6")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries loaded later may write to file descriptor 1 on their own (NCCL prints
# its version banner there when NCCL_DEBUG is set in the environment), so the process keeps a private duplicate of the
# real stdout for the result line and points descriptor 1 at stderr for everybody else.
_RESULT_OUT = None


def claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def build_map(n_keyframes, out=None):
    from uzliti_slam_b200 import synthetic as S
    t0 = time.time()
    kfs, pairs, poses = S.make_map(n_keyframes, n_features=N_FEATURES, cluster=25, pool=1000, n_shared=600,
                                   k_candidates=K_CAND, cross_cluster=4, seed=4, out=out)
    log(f"[bench] synthetic map: {n_keyframes} keyframes, {len(pairs)} candidate pairs in {time.time() - t0:.1f}s")
    return kfs, pairs, poses


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on all host cores (processes, because the oracle replays glibc rand())
# ----------------------------------------------------------------------------------------------------
_G = {}


def _cpu_worker(args):
    lo, hi = args
    from oracle import binding as O
    kfs, pairs = _G["kfs"], _G["pairs"]
    cons = 0
    for i in range(lo, hi):
        a, b = pairs[i]
        r = O.estimate_edge([kfs[a]], [kfs[b]], want_debug=False)
        cons += r["consensus"]
    return hi - lo, cons


def cpu_pass(kfs, pairs, pool, ncores, chunk=8):
    """one bounded CPU pass over `pairs`; returns (pairs done, seconds)"""
    _G["kfs"], _G["pairs"] = kfs, pairs
    jobs = [(i, min(i + chunk, len(pairs))) for i in range(0, len(pairs), chunk)]
    t0 = time.perf_counter()
    done = sum(n for n, _ in pool.map(_cpu_worker, jobs))
    return done, time.perf_counter() - t0


def make_pool(kfs, pairs, ncores):
    import multiprocessing as mp
    _G["kfs"], _G["pairs"] = kfs, pairs
    from oracle import binding as O
    O.lib()
    return mp.get_context("fork").Pool(ncores)


def cpu_baseline(kfs, pairs, budget_s=12.0):
    """oracle port on every host core for ~budget_s seconds of a fixed pseudo-random subsample"""
    ncores = len(os.sched_getaffinity(0))
    rng = np.random.default_rng(1)
    sample = pairs[rng.permutation(len(pairs))]
    pool = make_pool(kfs, sample, ncores)
    try:
        n0 = ncores * 8
        cpu_pass(kfs, sample[:n0], pool, ncores)                       # warm-up (page-in, pool start)
        _G["pairs"] = sample
        done, secs, pos = 0, 0.0, 0
        n = ncores * 32
        while secs < budget_s and pos < len(sample):
            jobs = [(i, min(i + 8, pos + n, len(sample))) for i in range(pos, min(pos + n, len(sample)), 8)]
            t0 = time.perf_counter()
            done += sum(k for k, _ in pool.map(_cpu_worker, jobs))
            secs += time.perf_counter() - t0
            pos += n
    finally:
        pool.close()
        pool.join()
    val = done / secs
    extra = {}
    try:   # context only: the OpenCV matcher the reference links, single thread, same 1000x1000 shape
        import cv2
        cv2.setNumThreads(1)
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        a, b = sample[0]
        t0 = time.perf_counter()
        for _ in range(5):
            bf.knnMatch(kfs[b]["desc"], kfs[a]["desc"], k=2)
        extra["cv2_knnmatch_ms_1thread"] = round((time.perf_counter() - t0) / 5 * 1e3, 2)
        extra["cv2_version"] = cv2.__version__
    except Exception:
        pass
    return dict(value=round(val, 2), unit=UNIT, cores=ncores, kind="port",
                sample=f"{done} pairs of the same workload (fixed random subsample of the 200000), oracle port "
                       f"(oracle/uz_oracle.cpp, g++ -O3 x86-64-v3), one process per core, {secs:.1f}s", **extra)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_kf = min(N_KEYFRAMES, 2000)      # the sample below never touches more keyframes than this
    kfs, pairs, _ = build_map(n_kf)
    ncores = len(os.sched_getaffinity(0))
    rng = np.random.default_rng(1)
    sample = pairs[rng.permutation(len(pairs))]
    per_step = ncores * 48
    pool = make_pool(kfs, sample, ncores)
    try:
        _G["pairs"] = sample
        pos = 0

        def step():
            nonlocal pos
            lo = pos % max(1, len(sample) - per_step)
            jobs = [(i, min(i + 8, lo + per_step)) for i in range(lo, lo + per_step, 8)]
            t0 = time.perf_counter()
            d = sum(k for k, _ in pool.map(_cpu_worker, jobs))
            pos += per_step
            return d, time.perf_counter() - t0
        for _ in range(args.warmup):
            step()
        done, secs = 0, 0.0
        for _ in range(args.steps):
            d, s = step()
            done += d
            secs += s
    finally:
        pool.close()
        pool.join()
    val = done / secs
    sample_desc = (f"each step = {per_step} pairs of the C4 workload (fixed random subsample), oracle port on "
                   f"{ncores} host processes")
    line = dict(impl="reference", metric=METRIC, value=round(val, 2), unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=round(secs / args.steps * 1e3, 3), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="u32", data="synthetic",
                config=dict(workload=WORKLOAD, l2="n/a (CPU)", sample=sample_desc),
                cpu_baseline=dict(value=round(val, 2), unit=UNIT, cores=ncores, kind="port", sample=sample_desc),
                e2e=dict(value=round(val, 2), unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    emit(line)


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region.  NVML is polled from a thread every 10 ms
    (started before the warm-up so the first sample never lands after the region); `mark()` brackets the
    timed region and only samples taken inside it are reported.  nvidia-smi is the fallback when NVML is absent."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.samples = []          # (t, sm_mhz, max_mhz, power_w, reasons-bitmask)
        self.t_begin = self.t_end = None
        self._stop = False
        self._thread = None
        self.source = None

    def _nvml_loop(self):
        import pynvml as N
        N.nvmlInit()
        # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
        idx = self.device
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                idx = int(vis.split(",")[self.device])
            except Exception:
                idx = self.device
        h = N.nvmlDeviceGetHandleByIndex(idx)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        self.source = "nvml"
        while not self._stop:
            try:
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                pw = N.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = N.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(N, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.perf_counter(), float(sm), float(mx), pw, int(rs)))
            except Exception:
                pass
            time.sleep(0.01)

    def _smi_loop(self):
        self.source = "nvidia-smi"
        proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                 "-i", str(self.device), "-lms", "50"], stdout=subprocess.PIPE, text=True)
        self._proc = proc
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        bits = (0x8, 0x40, 0x20, 0x4)
        for ln in proc.stdout:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                mask = 0
                for b, v in zip(bits, f[5:9]):
                    if v.lower().startswith("active"):
                        mask |= b
                self.samples.append((time.perf_counter(), float(f[1]), float(f[2]), float(f[3]), mask))
            except ValueError:
                continue
            if self._stop:
                break
        proc.terminate()

    def start(self):
        def run():
            try:
                self._nvml_loop()
            except Exception:
                try:
                    self._smi_loop()
                except Exception:
                    self.source = None
        self._thread = threading.Thread(target=run, daemon=True)
        self._thread.start()
        t0 = time.time()
        while not self.samples and time.time() - t0 < 5.0:      # first sample before anything is timed
            time.sleep(0.01)

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        self._stop = True
        if getattr(self, "_proc", None):
            self._proc.terminate()
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"], source=self.source)
        inside = [s for s in self.samples if self.t_begin is not None and self.t_begin <= s[0] <= (self.t_end or 1e300)]
        window = "timed region"
        if not inside:                       # region shorter than one sampling period: nearest samples around it
            window = "nearest samples around the timed region"
            mid = 0.5 * ((self.t_begin or 0) + (self.t_end or 0))
            inside = sorted(self.samples, key=lambda s: abs(s[0] - mid))[:3]
        mask = 0
        for s in inside:
            mask |= s[4]
        reasons = [n for n, b in (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20),
                                  ("hw_thermal_slowdown", 0x40), ("hw_power_brake", 0x80)) if mask & b]
        return dict(sm_mhz=float(np.median([s[1] for s in inside])), sm_max_mhz=float(max(s[2] for s in inside)),
                    power_w_max=float(max(s[3] for s in inside)), samples=len(inside), window=window,
                    source=self.source, reasons=reasons)


def measure_places(est, handles, kfs, my_pairs=None, cpu_keyframes=1500):
    """SURVEY 8f-1 (the step before the path): searchAndAddPlace for the whole map in one batch on the device vs the
    sequential oracle restatement of LshSetRecognizer on one host core (bounded to the first `cpu_keyframes`)."""
    from oracle import binding as O
    n = len(handles)
    stamps = (np.arange(n, dtype=np.int64) * 10_000_000_000)
    est.setPlaceConfig(T=2.0, k_nearest_neighbors=20)
    best = None
    for rep in range(3):
        est.clearPlaces()
        t0 = time.perf_counter()
        pairs = est.searchAndAddPlaces(handles, stamps)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, est.places_last_timing())
    dt, tm = best
    cluster = 25
    true = (pairs[:, 0] // cluster) == (pairs[:, 1] // cluster)          # handles are dense and in map order
    entries = n * N_FEATURES * 8
    P = O.Places(T=2.0, k=20)
    m = min(cpu_keyframes, n)
    t0 = time.perf_counter()
    want = [P.search_and_add(int(handles[i]), int(stamps[i]), [kfs[i]]) for i in range(m)]
    cpu_dt = time.perf_counter() - t0
    want = np.concatenate([w for w in want if len(w)] + [np.zeros((0, 2), np.int64)])
    got_head = pairs[pairs[:, 1] < handles[m - 1] + 1] if m < n else pairs
    same = bool(np.array_equal(got_head.astype(np.int64), want))
    return dict(what="searchAndAddPlace x %d keyframes (LSH-bucket voting, T=2, k=20), one batch" % n,
                keyframes_per_s=round(n / dt, 1), wall_ms=round(dt * 1e3, 2), insert_ms=round(tm["insert_ms"], 3),
                vote_ms=round(tm["vote_ms"], 3), select_ms=round(tm["select_ms"], 3), bucket_entries=entries,
                pairs=int(len(pairs)), true_pair_frac=round(float(true.mean()), 4) if len(pairs) else None,
                cpu_oracle=dict(keyframes=m, keyframes_per_s=round(m / cpu_dt, 1), cores=1, kind="port",
                                same_pairs_as_gpu=same))


def measure_other_configs(est, handles, kfs):
    """The other BASELINE.json configs, each as one short measurement beside the headline (rank 0, not part of the timed
    region; parity for these shapes is in tests/): C1 single pair latency, C3 one query vs 1000 candidates, C5 rigs of
    4 x 2000 features with 1000 hypotheses.  CPU figures are the single-thread oracle port on the same inputs."""
    from oracle import binding as O
    from uzliti_slam_b200 import synthetic as S
    out = {}

    # C1: one pair, 500 ORB-256 features each, through the store (handles) and from host buffers
    f, t, _ = S.make_pair(500, seed=1)
    h = est.add_keyframes([f, t])
    lat = []
    for _ in range(60):
        t0 = time.perf_counter()
        r = est.estimateEdges(h[:1], h[1:2])
        lat.append(time.perf_counter() - t0)
    lat_host = []
    for _ in range(60):
        t0 = time.perf_counter()
        est.estimateEdgesHost([([f], [t])])
        lat_host.append(time.perf_counter() - t0)
    cpu = []
    for _ in range(5):
        t0 = time.perf_counter()
        o = O.estimate_edge([f], [t])
        cpu.append(time.perf_counter() - t0)
    out["C1_single_pair_500"] = dict(latency_us_store=round(float(np.median(lat[10:])) * 1e6, 1),
                                     latency_us_host_buffers=round(float(np.median(lat_host[10:])) * 1e6, 1),
                                     cpu_port_1thread_us=round(float(np.median(cpu)) * 1e6, 1),
                                     same_consensus_as_cpu=bool(int(r[0]["consensus"]) == int(o["consensus"])))
    for x in h:
        est.remove_keyframe(int(x))

    # the online case: one new keyframe against its visual-odometry predecessor + 20 loop-closure candidates (N = 1000),
    # a launch far too small to fill the chip unless the train rows are cut into segments as well
    qo = np.full(21, handles[30], dtype=handles.dtype)
    co = handles[5:26]
    lat_o = []
    for _ in range(60):
        t0 = time.perf_counter()
        est.estimateEdges(co, qo)
        lat_o.append(time.perf_counter() - t0)
    out["online_21_pairs"] = dict(latency_us=round(float(np.median(lat_o[10:])) * 1e6, 1),
                                  note="1 keyframe vs 21 stored keyframes (2.1e7 compares + 21 RANSACs), records on the host")

    # C3: one query keyframe against 1000 candidates of the resident map (N = 1000)
    q = np.full(1000, handles[0], dtype=handles.dtype)
    cand = handles[1:1001]
    est.estimateEdges(cand, q)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        est.estimateEdges(cand, q)
    dt = (time.perf_counter() - t0) / reps
    out["C3_one_query_vs_1000"] = dict(ms_per_query=round(dt * 1e3, 3), edges_per_s=round(1000 / dt, 1),
                                       note="1e9 descriptor compares + 1000 RANSACs per query keyframe, records back on the host")

    # the same C3 batch with the opt-in cross-check (uz_params.cross_check; north star: "ratio test and cross-check"):
    # column minima tracked inside the match kernel
    est.enable_timers(True)
    ms = {}
    for cross in (0, 1):
        est.setConfig(cross_check=cross)
        est.estimateEdges(cand, q)
        est.reset_timers()
        for _ in range(reps):
            est.estimateEdges(cand, q)
        ms[cross] = est.get_timers()["match_ms"] / reps
    est.setConfig(cross_check=0)
    est.enable_timers(False)
    out["C3_cross_check"] = dict(match_ms_plain=round(ms[0], 3), match_ms_cross_check=round(ms[1], 3),
                                 cost_factor=round(ms[1] / ms[0], 3),
                                 note="fused form: per-train-row minima in the same pass (a second, reversed matching costs 2.0 x)")

    # C5: rig keyframes (4 cameras x 2000 features), 4 same-frame matchings of 2000 x 2000 per pair, 1000 hypotheses
    rigs_f, rigs_t = [], []
    for p in range(8):
        cf, ct = [], []
        for cam in range(4):
            a, b, _ = S.make_pair(2000, seed=900 + 10 * p + cam, rho=0.2 + 0.1 * ((cam + p) % 4), sensor_frame=cam)
            cf.append(a); ct.append(b)
        rigs_f.append(cf); rigs_t.append(ct)
    hf = est.add_keyframes(rigs_f)
    ht = est.add_keyframes(rigs_t)
    n5 = 592
    pf, pt = np.resize(hf, n5), np.resize(ht, n5)
    est.setConfig(ransac_iterations=1000)
    try:
        est.estimateEdges(pf, pt)
        est.enable_timers(True)
        est.reset_timers()
        t0 = time.perf_counter()
        for _ in range(3):
            r5 = est.estimateEdges(pf, pt)
        dt = (time.perf_counter() - t0) / 3
        tm = est.get_timers()
        est.enable_timers(False)
        t0 = time.perf_counter()
        o5 = O.estimate_edge(rigs_f[0], rigs_t[0], iterations=1000)
        cpu5 = time.perf_counter() - t0
    finally:
        est.setConfig(ransac_iterations=100)
    out["C5_rig_4x2000_1000hyp"] = dict(pairs=n5, edges_per_s=round(n5 / dt, 1),
                                        knn2_gcmp_per_s=round(tm["compares"] / (tm["match_ms"] * 1e-3) * 1e-9, 1),
                                        knn2_ms=round(tm["match_ms"] / 3, 3), solve_ms=round(tm["solve_ms"] / 3, 3),
                                        cpu_port_1thread_edges_per_s=round(1.0 / cpu5, 2),
                                        same_consensus_as_cpu=bool(int(r5[0]["consensus"]) == int(o5["consensus"])))
    for x in list(hf) + list(ht):
        est.remove_keyframe(int(x))

    # BRISK / FREAK rows (64 bytes; cv::BRISK is FeatureExtractionCore's default, feature_extraction_core.cpp:46-49): the same
    # loop-closure batch shape on a small map of 512-bit descriptors, through knn2_wide_kernel (8 POPC per compare)
    kw, pw, _ = S.make_map(400, n_features=1000, k_candidates=20, seed=77, desc_bytes=64)
    hw = est.add_keyframes(kw)
    est.estimateEdges(hw[pw[:, 0]], hw[pw[:, 1]])
    est.enable_timers(True)
    est.reset_timers()
    t0 = time.perf_counter()
    for _ in range(3):
        rw = est.estimateEdges(hw[pw[:, 0]], hw[pw[:, 1]])
    dt = (time.perf_counter() - t0) / 3
    tm = est.get_timers()
    est.enable_timers(False)
    t0 = time.perf_counter()
    ow = O.estimate_edge([kw[pw[0, 0]]], [kw[pw[0, 1]]])
    cpuw = time.perf_counter() - t0
    gcmp = tm["compares"] / (tm["match_ms"] * 1e-3) * 1e-9
    popc = est.microbench(0)
    out["BRISK512_loop_closure"] = dict(pairs=int(len(pw)), edges_per_s=round(len(pw) / dt, 1),
                                        knn2_wide_gcmp512_per_s=round(gcmp, 1), knn2_wide_ms=round(tm["match_ms"] / 3, 3),
                                        frac_of_popc_ceiling=round(gcmp * 8 / popc, 4),
                                        popc_per_compare=8, lop3_per_compare=26,
                                        cpu_port_1thread_edges_per_s=round(1.0 / cpuw, 2),
                                        same_consensus_as_cpu=bool(int(rw[0]["consensus"]) == int(ow["consensus"])))
    for x in hw:
        est.remove_keyframe(int(x))
    return out


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            log(f"[bench] --gpus {args.gpus} needs torchrun (one process per GPU); re-launching")
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"WORLD_SIZE={world} does not match --gpus {args.gpus}")

    n_kf = args.keyframes
    pairs_per_gpu = args.pairs_per_gpu

    # ---- data (host, pageable first: the CPU baseline forks before CUDA is touched) ----------------
    desc = np.empty((n_kf, N_FEATURES, 32), np.uint8)
    pos = np.empty((n_kf, N_FEATURES, 3), np.float64)
    valid = np.empty((n_kf, N_FEATURES), np.uint8)
    kfs, pairs, poses = build_map(n_kf, out=(desc, pos, valid))
    total_pairs = len(pairs)
    # rank r owns a contiguous chunk of the (from-sorted) pair list; cycle if the map is smaller than the job
    from uzliti_slam_b200.sharding import shard_bounds
    lo, hi = shard_bounds(world * pairs_per_gpu, world, rank)
    sel = np.arange(lo, hi) % total_pairs
    my_pairs = pairs[sel]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(kfs, pairs, budget_s=args.cpu_seconds)
        log(f"[bench] cpu_baseline: {cpu['value']} {UNIT} on {cpu['cores']} cores")

    import torch
    import torch.distributed as dist
    from uzliti_slam_b200 import EdgeEstimator
    from uzliti_slam_b200.binding import RESULT_DTYPE
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # pinned host copy of the map (the e2e leg DMAs straight out of it)
    t_desc = torch.from_numpy(desc).pin_memory()
    t_pos = torch.from_numpy(pos).pin_memory()
    t_valid = torch.from_numpy(valid).pin_memory()
    pd, pp, pv = t_desc.numpy(), t_pos.numpy(), t_valid.numpy()
    pinned_kfs = [dict(kf, desc=pd[i], pos=pp[i], valid=pv[i]) for i, kf in enumerate(kfs)]

    est = EdgeEstimator(local_rank)
    # a real (non-default) torch stream: the library launches on it, torch events and NCCL see the same queue
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    est.set_stream(stream.cuda_stream)

    # integer-pipe peaks, measured here and now (roofline denominators)
    peaks = {name: est.microbench(op) for op, name in enumerate(["popc", "lop3", "imad", "vimnmx"])}

    t0 = time.time()
    handles = est.add_keyframes(pinned_kfs)
    torch.cuda.synchronize()
    log(f"[bench] rank {rank}: store resident, {est.store_bytes() / 1e6:.0f} MB in {time.time() - t0:.2f}s")
    hf = np.ascontiguousarray(handles[my_pairs[:, 0]])
    ht = np.ascontiguousarray(handles[my_pairs[:, 1]])
    rec = RESULT_DTYPE.itemsize
    res_local = torch.empty(pairs_per_gpu * rec, dtype=torch.uint8, device=dev)
    res_all = torch.empty(world * pairs_per_gpu * rec, dtype=torch.uint8, device=dev) if world > 1 else res_local

    def step():
        est.estimateEdgesDevice(hf, ht, res_local.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(res_all, res_local)      # per-pair best edges over NCCL/NVLink

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    est.enable_timers(True)
    est.reset_timers()
    launches0 = est.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = est.launch_count() - launches0
    tm = est.get_timers()
    est.enable_timers(False)
    if world > 1:
        t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
        launches = int(t[1])
    value = world * pairs_per_gpu * args.steps / (ms * 1e-3)

    # sanity on the last step's records (not timed): true candidates found, false ones rejected
    out = np.frombuffer(res_local.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)
    same = (my_pairs[:, 0] // 25) == (my_pairs[:, 1] // 25)
    sanity = dict(ok_frac=float((out["ok"] == 1).mean()), median_consensus_true=float(np.median(out["consensus"][same])),
                  max_consensus_false=int(out["consensus"][~same].max()) if (~same).any() else 0,
                  mean_matches=float(out["n_matches"].mean()))

    # ---- e2e: the same batch from pinned host buffers through uz_estimate_edges_host ------------------
    prep = est.prepare_host_pairs([([pinned_kfs[a]], [pinned_kfs[b]]) for a, b in my_pairs])
    uniq = np.unique(my_pairs)
    h2d = int(len(uniq)) * N_FEATURES * (32 + 24 + 1)
    d2h = pairs_per_gpu * rec
    e2e_steps = max(2, min(args.steps, 5))
    est.estimateEdgesHostPrepared(prep)
    est.estimateEdgesHostPrepared(prep)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        est.estimateEdgesHostPrepared(prep)       # blocks until the edge records are back in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_val = world * pairs_per_gpu * e2e_steps / e2e_s
    assert prep["res"].tobytes() == out.tobytes(), "host path and store path disagree"

    # ---- the two kernels each on their own (not timed, rank 0): in the timed region the solve grid runs BESIDE the
    # match kernel (streaming form), so the live spans above overlap; this pass gives the un-shared durations
    alone = None
    if rank == 0:
        est.set_stream_solve(0)
        est.estimateEdgesDevice(hf, ht, res_local.data_ptr())      # (no collective here: rank 0 only)
        est.enable_timers(True)
        est.reset_timers()
        for _ in range(3):
            est.estimateEdgesDevice(hf, ht, res_local.data_ptr())
        ta = est.get_timers()
        est.enable_timers(False)
        est.set_stream_solve(STREAM_SOLVE)
        alone = dict(knn2_ms=ta["match_ms"] / max(ta["match_launches"], 1), solve_ms=ta["solve_ms"] / max(ta["solve_launches"], 1))
    barrier()

    places = None
    if rank == 0 and not args.no_places:
        places = measure_places(est, handles, kfs, my_pairs=None)

    others = None
    if rank == 0 and not args.no_extras:
        others = measure_other_configs(est, handles, kfs)

    if rank == 0:
        cmp_per_launch = tm["compares"] / max(tm["match_launches"], 1)
        knn_ms = tm["match_ms"] / max(tm["match_launches"], 1)
        solve_ms = tm["solve_ms"] / max(tm["solve_launches"], 1)
        gcmp = cmp_per_launch / (knn_ms * 1e-3) * 1e-9
        # per compare the kernel issues 13 LOP3 + 1.25 VIMNMX(3).U16x2 on the ALU pipe, 4 POPC on the XU pipe and
        # 4 IMAD on the FMA pipe (SASS of knn2_kernel<256,2,true,true>, profiles/); ncu charges a VIMNMX one ALU slot
        alu_limit = peaks["lop3"] / 14.25
        popc_limit = peaks["popc"] / 4.0
        peak = min(alu_limit, popc_limit)
        gcmp_alone = cmp_per_launch / (alone["knn2_ms"] * 1e-3) * 1e-9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "knn2_ncu_summary.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        alg_bytes = pairs_per_gpu * (2 * N_FEATURES * 32 + N_FEATURES * 8)
        mp = {}
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = mp.get("hbm_gbs", 6650.0)
        line = dict(
            metric=METRIC, value=round(value, 1), unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=round(ms / args.steps, 3), higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="u32", data="synthetic",
            config=dict(workload=WORKLOAD, keyframes=n_kf, features=N_FEATURES, pairs_per_gpu=pairs_per_gpu,
                        l2="inputs larger than L2: every step streams the 10000-keyframe store "
                           f"({est.store_bytes() / 1e6:.0f} MB resident per GPU) through the match kernel; no flush",
                        parallelism=f"pair-list sharding x{world}, store replicated, all-gather of 176 B edge records"),
            g_descriptor_cmp_per_s=round(world * cmp_per_launch / (ms / args.steps * 1e-3) * 1e-9, 2),
            roofline=dict(bound="int", kernel="knn2_kernel", achieved=round(gcmp, 2), peak=round(peak, 2), unit="Gcmp/s",
                          frac=round(gcmp / peak, 4), traffic=traffic,
                          peak_source="min(ALU, POPC) limit of the kernel's own SASS mix (13 LOP3 + 1.25 VIMNMX.U16x2 on the "
                                      "ALU pipe, 4 POPC on the XU pipe per 256-bit compare) from pipe rates microbenchmarked "
                                      "in this run; POPC binds",
                          live_span="knn2_ms_per_launch is the kernel's span in the timed region, where the persistent solve "
                                    "grid shares the SMs with it (streaming form); 'alone' is the same launch with the "
                                    "solve behind it (uz_set_stream_solve(0)), measured after the timed region",
                          alone=dict(knn2_ms_per_launch=round(alone["knn2_ms"], 3), solve_ms_per_launch=round(alone["solve_ms"], 3),
                                     achieved=round(gcmp_alone, 2), frac=round(gcmp_alone / peak, 4)),
                          stream_solve_ctas_per_sm=STREAM_SOLVE,
                          solve=dict(kernel="solve_kernel (one CTA per pair, alone)", bound="latency/issue, not HBM (reported as the north star asks)",
                                     algorithmic_bytes_per_pair=int(N_FEATURES * 8 + sanity["mean_matches"] * 48 + rec),
                                     achieved_gbs=round(pairs_per_gpu * (N_FEATURES * 8 + sanity["mean_matches"] * 48 + rec)
                                                        / (alone["solve_ms"] * 1e-3) * 1e-9, 2),
                                     frac_of_hbm=round(pairs_per_gpu * (N_FEATURES * 8 + sanity["mean_matches"] * 48 + rec)
                                                       / (alone["solve_ms"] * 1e-3) * 1e-9 / hbm_peak, 5)),
                          pipe_peaks_gops={k: round(v, 1) for k, v in peaks.items()},
                          textbook_peak_8popc=round(peaks["popc"] / 8.0, 2), frac_of_textbook=round(gcmp / (peaks["popc"] / 8.0), 4),
                          knn2_ms_per_launch=round(knn_ms, 3), solve_ms_per_launch=round(solve_ms, 3),
                          compares_per_launch=int(cmp_per_launch),
                          hbm=dict(achieved_gbs=round(alg_bytes / (knn_ms * 1e-3) * 1e-9, 2), peak_gbs=hbm_peak,
                                   frac=round(alg_bytes / (knn_ms * 1e-3) * 1e-9 / hbm_peak, 5),
                                   peak_source="MEASURED_PEAKS.json" if mp else "fallback")),
            cpu_baseline=cpu,
            e2e=dict(value=round(e2e_val, 1), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=e2e_steps,
                     api="uz_estimate_edges_host (pinned host FeatureData in, host edge records out)"),
            gpu_launches=launches, clocks=clocks, sanity=sanity, candidate_generation=places, other_configs=others)
        emit(line)
    est.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--keyframes", type=int, default=N_KEYFRAMES)
    ap.add_argument("--pairs-per-gpu", type=int, default=PAIRS_PER_GPU)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-places", action="store_true", help="skip the candidate-generation (8f-1) measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the short C1/C3/C5 measurements")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
