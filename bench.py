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
    # the splitmix64 generator that include/uz_synth.h mirrors byte for byte: a C++ host builds the same map
    from uzliti_slam_b200 import synth_splitmix as S
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
    recs = []
    for i in range(lo, hi):
        a, b = pairs[i]
        r = O.estimate_edge([kfs[a]], [kfs[b]], want_debug=False)
        recs.append((int(bool(r["ok"])), int(r["n_ratio_matches"]), int(r["n_matches"]), int(r["consensus"]),
                     np.asarray(r["T"], np.float64).tobytes()))
    return hi - lo, recs


def reference_faithful_1thread(kfs, sample, budget_s=4.0):
    """the reference's own execution model - ONE thread per estimator
    (transformation_estimator.cpp:26), OpenCV's BFMatcher for the matching (the call of :38,:58), then ratio test, depth
    filter, sort, gather and the oracle's RANSAC (PCL/Eigen restated), minus the reference's 1 ms sleep per pair."""
    try:
        import cv2
    except Exception:
        return None
    from oracle import binding as O
    cv2.setNumThreads(1)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    done, t0, checked = 0, time.perf_counter(), 0
    for a, b in sample:
        F, T = kfs[a], kfs[b]
        knn = bf.knnMatch(T["desc"], F["desc"], k=2)                                     # query = to, train = from (:58)
        m = np.array([(x[0].queryIdx, x[0].trainIdx, x[0].distance, x[1].distance) for x in knn if len(x) == 2], np.float64)
        keep = m[:, 2] < 0.99 * m[:, 3]                                                  # :65-71
        q, t, d = m[keep, 0].astype(np.int64), m[keep, 1].astype(np.int64), m[keep, 2]
        v = (T["valid"][q] != 0) & (F["valid"][t] != 0)                                  # :103-112
        q, t, d = q[v], t[v], d[v]
        o = np.lexsort((q, d))                                                           # :114 in the (distance, queryIdx) order
        P, Q = T["pos"][q[o]], F["pos"][t[o]]                                            # :118-124
        r = O.estimate_svd(P, Q, 0.1, 100, 0.6) if len(o) >= 3 else None                 # :130
        if checked < 3 and r is not None:                                                # the composed pipeline IS the oracle's
            full = O.estimate_edge([F], [T], want_debug=False)
            assert full["consensus"] == r["consensus"] and full["n_matches"] == len(o)
            checked += 1
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    secs = time.perf_counter() - t0
    return dict(value=round(done / secs, 2), unit=UNIT, cores=1, pairs=done, seconds=round(secs, 2),
                what="cv2.BFMatcher(NORM_HAMMING).knnMatch (OpenCV %s, 1 thread) + ratio/depth filter/sort + oracle RANSAC, one pair "
                     "after the other (the reference's single estimator thread without its 1 ms sleep)" % cv2.__version__)


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
        records = []
        while secs < budget_s and pos < len(sample):
            jobs = [(i, min(i + 8, pos + n, len(sample))) for i in range(pos, min(pos + n, len(sample)), 8)]
            t0 = time.perf_counter()
            for k, recs in pool.map(_cpu_worker, jobs):
                done += k
                records += recs
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
    faithful = reference_faithful_1thread(kfs, sample[:4000])
    if faithful:
        extra["reference_faithful_1thread"] = faithful
    out = dict(value=round(val, 2), unit=UNIT, cores=ncores, kind="port",
               sample=f"{done} pairs of the same workload (fixed random subsample of the 200000), oracle port "
                      f"(oracle/uz_oracle.cpp, g++ -O3 x86-64-v3), one process per core, {secs:.1f}s", **extra)
    return out, sample[:len(records)], records


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kfs, pairs, _ = build_map(N_KEYFRAMES)      # the same 10000-keyframe map and pair list as the GPU arm
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
    sample_desc = (f"each step = {per_step} pairs of the C4 workload (fixed random subsample of the same 200000 pairs on the "
                   f"same {N_KEYFRAMES}-keyframe map; per-pair work is identical), oracle port on {ncores} host processes")
    line = dict(impl="reference", metric=METRIC, value=round(val, 2), unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=round(secs / args.steps * 1e3, 3), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="u32", data="synthetic",
                config=dict(workload=WORKLOAD, keyframes=N_KEYFRAMES, features=N_FEATURES, l2="n/a (CPU)", sample=sample_desc),
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
    # HBM view of the two table kernels (random 16 B slots + 8 B nodes: bound by DRAM sectors, not bytes): algorithmic bytes per
    # bucket entry = 4 B key + 16 B slot + 8 B node; `traffic` = what ncu saw (profiles/places_ncu_summary.json)
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        mp = {}
    hbm_peak = mp.get("hbm_gbs", 6650.0)
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "places_ncu_summary.json")))
    except Exception:
        ncu = {}
    alg = entries * 28.0
    roof = {}
    for name, ms_k in (("insert", tm["insert_ms"]), ("vote", tm["vote_ms"])):
        gbs = alg / (ms_k * 1e-3) * 1e-9 if ms_k > 0 else 0.0
        traffic = ncu.get(name + "_dram_bytes_per_launch") if n == N_KEYFRAMES else None
        roof[name] = dict(bound="hbm", achieved=round(gbs, 1), peak=hbm_peak, unit="GB/s", frac=round(gbs / hbm_peak, 4),
                          algorithmic_bytes=int(alg), traffic=traffic,
                          traffic_gbs=round(traffic / (ms_k * 1e-3) * 1e-9, 1) if traffic and ms_k > 0 else None,
                          note="random 32 B sectors: traffic / algorithmic bytes is the sector overfetch of an open-addressing table")
    return dict(what="searchAndAddPlace x %d keyframes (LSH-bucket voting, T=2, k=20), one batch" % n,
                keyframes_per_s=round(n / dt, 1), wall_ms=round(dt * 1e3, 2), insert_ms=round(tm["insert_ms"], 3),
                vote_ms=round(tm["vote_ms"], 3), select_ms=round(tm["select_ms"], 3), bucket_entries=entries,
                pairs=int(len(pairs)), true_pair_frac=round(float(true.mean()), 4) if len(pairs) else None, roofline=roof,
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
                                 note="256-bit rows on the tensor cores: the matchings run a second time with the roles swapped (the integer-pipe "
                                      "kernels fuse the per-train-row minima into one pass at 1.15 x, but are 5 x slower to begin with)")

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
    # loop-closure batch shape on a small map of 512-bit descriptors, through knn2_mmaf_kernel<true> (tensor cores, 4-bit
    # operands, K = 512 in eight instructions; UZ_MATCH_MMA_WIDE=2 selects knn2_mmaw_kernel on two int8 planes,
    # UZ_MATCH_MMA_WIDE=0 knn2_wide_kernel on the integer pipes, 8 POPC per compare)
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
    try:
        bf16_burst = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops", 1590.0)
    except Exception:
        bf16_burst = 1590.0
    out["BRISK512_loop_closure"] = dict(pairs=int(len(pw)), edges_per_s=round(len(pw) / dt, 1),
                                        kernel=("knn2_mmaf_kernel<true> (tcgen05.mma kind::mxf4.block_scale, 128x240x64 on 4-bit operands, "
                                                "K = 512 in eight instructions)" if os.environ.get("UZ_MATCH_MMA_WIDE", "1") == "1" else
                                                "UZ_MATCH_MMA_WIDE=" + os.environ["UZ_MATCH_MMA_WIDE"]),
                                        knn2_wide_gcmp512_per_s=round(gcmp, 1), knn2_wide_ms=round(tm["match_ms"] / 3, 3),
                                        tops=round(gcmp * 1024 * 1e-3, 1), frac_of_4x_bf16_burst=round(gcmp * 1024 * 1e-3 / (4 * bf16_burst), 4),
                                        x_integer_pipe_popc_ceiling=round(gcmp * 8 / popc, 3),
                                        cpu_port_1thread_edges_per_s=round(1.0 / cpuw, 2),
                                        same_consensus_as_cpu=bool(int(rw[0]["consensus"]) == int(ow["consensus"])))
    for x in hw:
        est.remove_keyframe(int(x))
    return out


def _new_estimator(device, **env):
    """a context with UZ_* knobs set for its creation only (they are read by uz_create)"""
    from uzliti_slam_b200 import EdgeEstimator
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return EdgeEstimator(device)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            log(f"[bench] --gpus {args.gpus} needs torchrun (one process per GPU); re-launching")
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
            # (descriptor 1 of this process already points at stderr: hand the children the REAL stdout for the result line)
            sys.exit(subprocess.call(cmd, stdout=_RESULT_OUT or sys.stdout))
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

    cpu, cpu_pairs, cpu_records = None, None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, cpu_pairs, cpu_records = cpu_baseline(kfs, pairs, budget_s=args.cpu_seconds)
        log(f"[bench] cpu_baseline: {cpu['value']} {UNIT} on {cpu['cores']} cores, "
            f"single thread with cv2: {(cpu.get('reference_faithful_1thread') or {}).get('value')}")

    import torch
    import torch.distributed as dist
    from uzliti_slam_b200 import EdgeEstimator
    from uzliti_slam_b200.binding import RESULT_DTYPE
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")        # host-side waits that must not park a kernel on the GPUs

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

    # integer-pipe peaks, measured here and now (roofline denominators of the integer-pipe kernels)
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

    def timed(fn, steps, warm=1):
        """device time (ms, CUDA events on the launching stream) of `steps` calls, max over ranks"""
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        t = a.elapsed_time(b)
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt[0])
        return t / steps

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

    # C4 parity beyond the test suite's sample: every pair the CPU baseline solved (its records are kept) is solved again on
    # the GPU and compared field by field, transform bit for bit (untimed)
    if cpu_records:
        g = est.estimateEdges(handles[cpu_pairs[:, 0]], handles[cpu_pairs[:, 1]])
        bad = 0
        for r, (ok, nr, nm, cons, tb) in zip(g, cpu_records):
            if (int(r["ok"]), int(r["n_ratio_matches"]), int(r["n_matches"]), int(r["consensus"])) != (ok, nr, nm, cons) or \
                    (ok and r["T"].tobytes() != tb):
                bad += 1
        sanity["cpu_records_compared"] = len(cpu_records)
        sanity["cpu_records_equal"] = bad == 0
        sanity["cpu_records_differing"] = bad

    # N GPUs == 1 GPU, byte for byte: the head of the gathered batch, recomputed by rank 0 alone
    if world > 1:
        n_chk = min(2500, pairs_per_gpu)
        gathered = np.frombuffer(res_all.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)
        if rank == 0:
            last = world - 1
            llo, _ = shard_bounds(world * pairs_per_gpu, world, last)
            chk = np.concatenate([np.arange(0, n_chk), np.arange(llo, llo + n_chk)])      # rank 0's head and the last rank's head
            gp = pairs[chk % total_pairs]
            alone = est.estimateEdges(handles[gp[:, 0]], handles[gp[:, 1]])
            sanity["gathered_equals_single_gpu"] = bool(alone.tobytes() == gathered[chk].tobytes())
            sanity["gathered_pairs_compared"] = int(len(chk))

    # ---- e2e: the same batch from HOST buffers through uz_estimate_edges_host ---------------------------
    uniq = np.unique(my_pairs)
    h2d = int(len(uniq)) * N_FEATURES * (32 + 24 + 1)
    d2h = pairs_per_gpu * rec
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_leg(source):
        prep = est.prepare_host_pairs([([source[a]], [source[b]]) for a, b in my_pairs])
        est.estimateEdgesHostPrepared(prep)
        est.estimateEdgesHostPrepared(prep)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            est.estimateEdgesHostPrepared(prep)       # blocks until the edge records are back in host memory
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([secs], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t[0])
        assert prep["res"].tobytes() == out.tobytes(), "host path and store path disagree"
        return world * pairs_per_gpu * e2e_steps / secs

    e2e_val = e2e_leg(pinned_kfs)                      # pinned, device-mapped FeatureData: pulled by the gather kernel
    e2e_pageable = e2e_leg(kfs) if not args.no_extras else None    # plain malloc'ed arrays (cv::Mat / Eigen): pinned ring first

    # ---- the integer-pipe match kernel on the same batch (the north star's POPC design; UZ_MATCH_MMA=0) -------------
    int_pipe = None
    if rank == 0 and not args.no_extras:
        ep = _new_estimator(local_rank, UZ_MATCH_MMA=0, UZ_STREAM_SOLVE=0)
        ep.set_stream(stream.cuda_stream)
        hp = ep.add_keyframes(pinned_kfs)
        hpf, hpt = np.ascontiguousarray(hp[my_pairs[:, 0]]), np.ascontiguousarray(hp[my_pairs[:, 1]])
        ep.estimateEdgesDevice(hpf, hpt, res_local.data_ptr())
        ep.enable_timers(True)
        ep.reset_timers()
        for _ in range(3):
            ep.estimateEdgesDevice(hpf, hpt, res_local.data_ptr())
        tp = ep.get_timers()
        torch.cuda.synchronize()
        same_records = np.frombuffer(res_local.cpu().numpy().tobytes(), dtype=RESULT_DTYPE).tobytes() == out.tobytes()
        ep.close()
        kms = tp["match_ms"] / max(tp["match_launches"], 1)
        g = tp["compares"] / max(tp["match_launches"], 1) / (kms * 1e-3) * 1e-9
        from uzliti_slam_b200 import mix
        popc_limit, alu_limit = peaks["popc"] / mix.KNN2_POPC, peaks["lop3"] / (mix.KNN2_LOP3 + mix.KNN2_MINMAX)
        int_pipe = dict(kernel="knn2_kernel<256,2> (4 POPC + 13 LOP3 + 4 IMAD + 1.25 VIMNMX.U16x2 per 256-bit compare), solve behind it",
                        knn2_ms_per_launch=round(kms, 3), solve_ms_per_launch=round(tp["solve_ms"] / max(tp["solve_launches"], 1), 3),
                        achieved_gcmp_per_s=round(g, 2), popc_ceiling_gcmp_per_s=round(popc_limit, 2),
                        alu_ceiling_gcmp_per_s=round(alu_limit, 2), frac_of_popc_ceiling=round(g / min(popc_limit, alu_limit), 4),
                        textbook_8popc_ceiling=round(peaks["popc"] / 8.0, 2), frac_of_textbook=round(g / (peaks["popc"] / 8.0), 4),
                        pipe_peaks_gops={k: round(v, 1) for k, v in peaks.items()}, records_equal_tensor_core_path=bool(same_records))
    barrier()

    # ---- strong scaling and the latency-bound end (C3) at this N, through the same step (shards + all-gather) ----------
    def sharded(total, steps):
        """`total` pairs as one job over the ranks: every rank takes its contiguous shard in calls of at most 25000 pairs (the
        calls are asynchronous: the host prepares call k+1 while the GPU runs call k), then the records are all-gathered"""
        lo2, hi2 = shard_bounds(total, world, rank)
        idx = np.arange(lo2, hi2) % total_pairs
        f2 = np.ascontiguousarray(handles[pairs[idx, 0]])
        t2 = np.ascontiguousarray(handles[pairs[idx, 1]])
        per = (total + world - 1) // world
        loc = torch.empty(max(per, 1) * rec, dtype=torch.uint8, device=dev)
        allb = torch.empty(world * max(per, 1) * rec, dtype=torch.uint8, device=dev) if world > 1 else loc
        cuts = list(range(0, len(f2), 25000)) + [len(f2)]

        def fn():
            for a, b in zip(cuts[:-1], cuts[1:]):
                est.estimateEdgesDevice(f2[a:b], t2[a:b], loc.data_ptr() + a * rec)
            if world > 1:
                dist.all_gather_into_tensor(allb, loc)
        return timed(fn, steps)

    scaling_extra = None
    if not args.no_extras:
        strong_ms = sharded(200000, 2)
        c3_f = np.ascontiguousarray(handles[1:1001])
        c3_t = np.full(1000, handles[0], dtype=handles.dtype)
        lo3, hi3 = shard_bounds(1000, world, rank)
        per3 = (1000 + world - 1) // world
        loc3 = torch.empty(per3 * rec, dtype=torch.uint8, device=dev)
        all3 = torch.empty(world * per3 * rec, dtype=torch.uint8, device=dev) if world > 1 else loc3

        def c3():
            est.estimateEdgesDevice(c3_f[lo3:hi3], c3_t[lo3:hi3], loc3.data_ptr())
            if world > 1:
                dist.all_gather_into_tensor(all3, loc3)
        c3_ms = timed(c3, 20, warm=3)
        scaling_extra = dict(
            strong=dict(what="C4 as ONE job: 200000 pairs fixed, cut into %d contiguous shards, records all-gathered" % world,
                        ms=round(strong_ms, 3), edges_per_s=round(200000 / (strong_ms * 1e-3), 1)),
            c3=dict(what="C3: 1 query keyframe vs 1000 candidates (1000 pairs) cut into %d shards, records all-gathered: the "
                         "latency-bound end of the curve" % world,
                    ms=round(c3_ms, 4), edges_per_s=round(1000 / (c3_ms * 1e-3), 1)))

    # ---- the same job through ONE process: uz_group over all N devices (rank 0 drives, the other ranks wait on the host) ---
    group = None
    if world > 1 and not args.no_extras:
        if rank == 0:
            from uzliti_slam_b200 import GroupEstimator
            g = GroupEstimator(list(range(world)))
            t0 = time.perf_counter()
            hg = g.add_keyframes(pinned_kfs)
            t_store = time.perf_counter() - t0
            allp = pairs[np.arange(world * pairs_per_gpu) % total_pairs]
            gf, gt = np.ascontiguousarray(hg[allp[:, 0]]), np.ascontiguousarray(hg[allp[:, 1]])
            res_g = np.zeros(len(allp), RESULT_DTYPE)
            modes = {}
            for mode in (0, 1):
                g.set_gather(mode)
                g.estimateEdges(gf, gt, out=res_g)
                t0 = time.perf_counter()
                for _ in range(3):
                    g.estimateEdges(gf, gt, out=res_g)
                modes[mode] = (time.perf_counter() - t0) / 3
            dev_ms = float(g.last_timing().max())
            # records into device 0's memory: written by every device's solve kernel through the peer-mapped pointer (NVLink)
            buf_g = torch.empty(len(allp) * rec, dtype=torch.uint8, device=dev)
            g.estimateEdgesDevice(gf, gt, buf_g.data_ptr())
            t0 = time.perf_counter()
            for _ in range(3):
                g.estimateEdgesDevice(gf, gt, buf_g.data_ptr())
            peer_s = (time.perf_counter() - t0) / 3
            peer_ms = float(g.last_timing().max())
            peer_equal = bool(buf_g.cpu().numpy().tobytes() == res_g.tobytes())
            g.set_gather(1)
            c3g = []
            for _ in range(30):
                t0 = time.perf_counter()
                g.estimateEdges(c3_f.astype(np.int32) - handles[0] + hg[0], c3_t.astype(np.int32) - handles[0] + hg[0], out=res_g[:1000])
                c3g.append(time.perf_counter() - t0)
            n_chk = min(2500, pairs_per_gpu)
            chk = np.concatenate([np.arange(0, n_chk), np.arange(len(allp) - n_chk, len(allp))])
            g.estimateEdges(gf, gt, out=res_g)
            alone = est.estimateEdges(handles[allp[chk, 0]], handles[allp[chk, 1]])
            group = dict(what="uz_group_estimate_edges: ONE process, %d devices, replicated store pulled over NVLink, contiguous pair "
                              "shards, records delivered into ONE host array by one copy per device (no collective)" % world,
                         pairs=int(len(allp)), edges_per_s=round(len(allp) / modes[1], 1), ms=round(modes[1] * 1e3, 3),
                         device_ms_max=round(dev_ms, 3),
                         host_mapped_sink=dict(edges_per_s=round(len(allp) / modes[0], 1), ms=round(modes[0] * 1e3, 3),
                                               what="measured alternative: solve kernels write through a host-mapped pointer"),
                         peer_write=dict(edges_per_s=round(len(allp) / peer_s, 1), ms=round(peer_s * 1e3, 3), device_ms_max=round(peer_ms, 3),
                                         records_equal=peer_equal,
                                         what="uz_group_estimate_edges_device: records written into device 0's memory by every device's "
                                              "solve kernel over NVLink (gather fused into the solve)"),
                         store_replication_s=round(t_store, 3), c3_ms=round(float(np.median(c3g[5:])) * 1e3, 4),
                         records_equal_single_gpu=bool(alone.tobytes() == res_g[chk].tobytes()), pairs_compared=int(len(chk)))
            g.close()
        dist.barrier(group=cpu_group)
    barrier()

    places = None
    if rank == 0 and not args.no_places:
        places = measure_places(est, handles, kfs, my_pairs=None)

    others = None
    if rank == 0 and not args.no_extras:
        others = measure_other_configs(est, handles, kfs)

    adapter_queue = None
    if rank == 0 and world == 1 and not args.no_extras:
        # the same map and pair list through the C++ adapter's queue interface (the class a maintainer swaps in,
        # INTEGRATION.md): adapter/bench_adapter builds the map with include/uz_synth.h - the generator of build_map
        exe = os.path.join(ROOT, "adapter", "bench_adapter")
        if os.path.exists(exe):
            try:
                from uzliti_slam_b200 import synth_splitmix as SM
                out = subprocess.run([exe, str(n_kf), str(pairs_per_gpu), "3"], capture_output=True, text=True, timeout=300)
                adapter_queue = json.loads(out.stdout.strip().splitlines()[-1])
                adapter_queue["same_map_as_this_bench"] = bool(
                    adapter_queue.get("map_checksum_desc") == SM.checksum(np.stack([k["desc"] for k in kfs])))
            except Exception as e:  # noqa: BLE001
                adapter_queue = {"error": repr(e)[:200]}

    if rank == 0:
        cmp_per_launch = tm["compares"] / max(tm["match_launches"], 1)
        knn_ms = tm["match_ms"] / max(tm["match_launches"], 1)
        solve_ms = tm["solve_ms"] / max(tm["solve_launches"], 1)
        gcmp = cmp_per_launch / (knn_ms * 1e-3) * 1e-9
        mp = {}
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = mp.get("hbm_gbs", 6650.0)
        # tensor roofline: a 256-bit compare is a K = 256 dot product of 4-bit operands = 512 ops; the block-scaled fp4 peak is
        # 4 x the bf16 peak (same tcgen05 data path at four times the K per instruction); MEASURED_PEAKS.json holds the
        # measured bf16 figures
        bf16 = mp.get("bf16_tflops_sustained", 1400.0)
        bf16_burst = mp.get("bf16_tflops", 1590.0)
        from uzliti_slam_b200 import mix
        tops = gcmp * mix.MMA_OPS_PER_COMPARE * 1e-3
        traffic = None
        prof = os.path.join(ROOT, "profiles", "knn2_mma_ncu_summary.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        alg_bytes = pairs_per_gpu * (2 * N_FEATURES * 128 + N_FEATURES * 8)
        solve_bytes = int(N_FEATURES * 8 + sanity["mean_matches"] * 48 + rec)
        line = dict(
            metric=METRIC, value=round(value, 1), unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=round(ms / args.steps, 3), higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="e2m1 (exact dot products of +-4 four-bit operands, block scale 2, fp32 accumulate below 2^24) + u32 keys", data="synthetic",
            config=dict(workload=WORKLOAD, keyframes=n_kf, features=N_FEATURES, pairs_per_gpu=pairs_per_gpu,
                        l2="inputs larger than L2: every step streams the 10000-keyframe store "
                           f"({est.store_bytes() / 1e6:.0f} MB resident per GPU, 1.3 GB of it the 4-bit operand layout) through the "
                           "match kernel; no flush",
                        parallelism=f"pair-list sharding x{world}, store replicated, all-gather of 176 B edge records"),
            g_descriptor_cmp_per_s=round(world * cmp_per_launch / (ms / args.steps * 1e-3) * 1e-9, 2),
            roofline=dict(bound="tensor", kernel="knn2_mmaf_kernel (tcgen05.mma kind::mxf4.block_scale, 128x240x64 on 4-bit operands, TMEM "
                                                 "accumulators; the low half of the fp32 accumulator is the packed key)",
                          achieved=round(tops, 1), peak=round(4 * bf16_burst, 1), unit="TFLOP/s", frac=round(tops / (4 * bf16_burst), 4),
                          ops="one 4-bit multiply-add = 2 ops; a 256-bit compare = 512 ops",
                          traffic=traffic,
                          peak_source=("4 x bf16_tflops (burst) of MEASURED_PEAKS.json: block-scaled fp4 is the bf16 data path at four times the K "
                                       "per instruction (nominal 9 PFLOP/s dense); the match kernel runs in 3 ms bursts between solve "
                                       "launches that leave the tensor pipe idle, hence the burst figure.  scripts/mxf4_probe.cu measures "
                                       "141-155 clocks per 128x256x64 instruction against 128 nominal: the pipe itself tops out at ~0.87 of "
                                       "this peak" if mp else "fallback 4 x 1590 TF/s"),
                          frac_of_sustained_peak=round(tops / (4 * bf16), 4), sustained_peak=round(4 * bf16, 1),
                          frac_of_nominal_9000=round(tops / 9000.0, 4),
                          frac_of_int8_peak=round(tops / (2 * bf16_burst), 4),
                          achieved_gcmp_per_s=round(gcmp, 2), knn2_ms_per_launch=round(knn_ms, 3), solve_ms_per_launch=round(solve_ms, 3),
                          compares_per_launch=int(cmp_per_launch),
                          padding="1000 x 1000 matchings run as 4 items x (2 x 128) query rows x (4 x 240 + 48) train rows (96.9 % of the "
                                  "tile area is real compares), 4 + 1 instructions per tile (a kind::f8f6f4 K = 32 instruction starts the "
                                  "accumulator at the key offset and takes as long as one of the four): 77.5 % of the issued MMA time is "
                                  "counted",
                          hbm=dict(achieved_gbs=round(alg_bytes / (knn_ms * 1e-3) * 1e-9, 2), peak_gbs=hbm_peak,
                                   frac=round(alg_bytes / (knn_ms * 1e-3) * 1e-9 / hbm_peak, 5),
                                   algorithmic_bytes_per_launch=int(alg_bytes),
                                   peak_source="MEASURED_PEAKS.json" if mp else "fallback"),
                          solve=dict(kernel="solve_kernel (one CTA per pair, behind the match kernel)",
                                     bound="latency/issue, not HBM (reported as the north star asks)",
                                     algorithmic_bytes_per_pair=solve_bytes,
                                     achieved_gbs=round(pairs_per_gpu * solve_bytes / (solve_ms * 1e-3) * 1e-9, 2),
                                     frac_of_hbm=round(pairs_per_gpu * solve_bytes / (solve_ms * 1e-3) * 1e-9 / hbm_peak, 5)),
                          int_pipe_kernel=int_pipe),
            cpu_baseline=cpu,
            e2e=dict(value=round(e2e_val, 1), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=e2e_steps,
                     api="uz_estimate_edges_host (host FeatureData in, host edge records out)",
                     pinned=round(e2e_val, 1), pageable=round(e2e_pageable, 1) if e2e_pageable else None,
                     note="value = pinned, device-mapped host buffers (pulled over PCIe by the gather kernel); pageable = plain "
                          "malloc'ed arrays as cv::Mat / Eigen hold them (host threads pack them into a pinned ring first)"
                          + ("; at N > 1 every rank times its own host round trip (max over ranks), records stay in the rank's "
                             "host memory - the one-array form is `group`" if world > 1 else "")),
            gpu_launches=launches, clocks=clocks, sanity=sanity, scaling_extra=scaling_extra, group=group,
            candidate_generation=places, other_configs=others, adapter_queue=adapter_queue)
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
