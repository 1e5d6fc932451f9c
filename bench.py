#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Ligero hot path.

Metric (BASELINE.json): witness Fr-elements/sec (encode+commit).
Workload at N=1: BASELINE config 3 -- Ligero encode + SHA-256 Merkle commit of a synthetic
R = 2^22 rows x k = 256 message columns witness (n = 4k = 1024 codeword columns), BN254 Fr,
32 GiB resident in HBM, seed 3 (uniform canonical elements, finite_field_gmp.hpp:70-78).
One "step" = one full commit of the witness: every row Reed-Solomon encoded (iNTT_k -> NTT_4k),
every codeword column SHA-256 hashed over all rows in order, Merkle tree over the n leaves.

  value : elements/s with the witness already resident in HBM (device time, CUDA events)
  e2e   : same metric through the C-ABI call lgr_encode_commit_host with HOST (pinned) rows:
          H2D of every tile and D2H of the root inside the timed region
  N > 1 : one process per GPU, each commits its own R-row shard (weak scaling); the n leaf digests
          of every shard are all-gathered over NCCL and every rank builds the tree over the G*n
          leaves (DESIGN.md "Multi-GPU").  value = G*R*k / max-over-ranks device time.

`--impl reference` times the CPU oracle port (the reference has no CPU implementation of this path; what can be
built from its sources here is its shaders run scalar on the host, oracle/_ref -- a checker, not a baseline:
DESIGN.md "Oracle") on the host cores, on a bounded sample of the same workload.  Only that leg and the
cpu_baseline leg touch oracle/; the bench line's `parity_check` compares the CPU leg's root with the GPU's.
"""
import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "witness Fr-elements/sec (encode+commit)"
UNIT = "elements/s"
K = 256
LOG_ROWS = 22
SEED = 3
L2_BYTES = 126 * 1024 * 1024


def load_package():
    if "ligero_prover_b200" in sys.modules:
        return sys.modules["ligero_prover_b200"]
    path = os.path.join(ROOT, "ligero-prover_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("ligero_prover_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ligero_prover_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.device)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 0


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(target_seconds, k=K):
    """CPU oracle (oracle/liblgo.so, OpenMP over all host cores) on a bounded sample of the workload"""
    from oracle import lgo
    lgo.build()
    lgo.set_threads(len(os.sched_getaffinity(0)))          # torchrun exports OMP_NUM_THREADS=1
    cores = lgo.num_threads()
    t0 = time.perf_counter()
    lgo.encode_commit_synth(SEED, 1 << 11, k)              # calibration: 2^11 rows
    dt = max(time.perf_counter() - t0, 1e-4)
    rows = int((1 << 11) * target_seconds / dt)
    rows = max(1 << 11, min(rows, 1 << LOG_ROWS))
    rows = 1 << (rows.bit_length() - 1)
    t0 = time.perf_counter()
    _, nodes, _ = lgo.encode_commit_synth(SEED, rows, k)
    dt = time.perf_counter() - t0
    cpu_reference_rate.last_root = nodes[0].tobytes().hex()   # root of the first `rows` rows of the seed-3 matrix
    return rows * k / dt, cores, rows, dt


def run_reference(args, rank):
    if rank != 0:
        return
    per_step = max(2.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    vals = []
    cores = rows = 0
    for i in range(args.warmup + args.steps):
        v, cores, rows, dt = cpu_reference_rate(per_step)
        if i >= args.warmup:
            vals.append((v, dt))
    value = statistics.mean(v for v, _ in vals)
    ms = statistics.mean(dt for _, dt in vals) * 1e3
    sample = "oracle encode+commit of 2^%d rows x k=%d (seed %d) per step, OpenMP" % (rows.bit_length() - 1, K, SEED)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "ligero encode+commit, R=2^%d rows x k=%d (n=%d) per GPU, BN254 Fr, seed %d" % (LOG_ROWS, K, 4 * K, SEED),
                   "rows_per_gpu": 1 << LOG_ROWS, "k": K, "n": 4 * K, "sample_rows_per_step": rows,
                   "note": "the reference has no CPU path for these kernels; this is the CPU oracle port (pinned to the reference's shaders, tests/test_wgslref_cpu.py)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/*_ncu_full_summary.csv, written by tools/summarize_profiles.py; the capture profiles a launch of the same
    tile size as the bench's).  Returns (bytes or None, source)."""
    import csv, glob
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "*_ncu_full_summary.csv")), reverse=True):
        tot, seen = 0.0, 0
        for row in csv.reader(open(path)):
            if len(row) >= 5 and kernel in row[1] and row[2] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(row[3]) * unit.get(row[4], 1.0)
                seen += 1
        if seen == 2:
            return tot, os.path.relpath(path, os.path.dirname(os.path.abspath(__file__)))
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-rows", type=int, default=LOG_ROWS, help="rows per GPU = 2^log_rows (default: BASELINE config 3)")
    ap.add_argument("--k", type=int, default=K)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--layout", default="sharded", choices=["sharded", "exact"],
                    help="N>1: 'sharded' = per-rank commitments + one digest all-gather (north_star); 'exact' = bit-exact single root "
                         "via all-to-all of codeword column slabs (ligero-prover_b200/sharding.py)")
    ap.add_argument("--no-exact", action="store_true", help="skip the exact-layout / k=8192 legs (N>1: exact, exact_k8192; N=1: encode_commit_k8192)")
    ap.add_argument("--no-config5", action="store_true", help="N=8: skip the 2^26-row leg (BASELINE config 5)")
    ap.add_argument("--aux", action="store_true", help="also time the reference's default geometry k=8192 and the 2^20 NTT (extra keys)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; the hot path has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lgr = load_package()
    k, n = args.k, 4 * args.k
    R = 1 << args.log_rows
    dev = torch.device("cuda", local_rank)

    ex = lgr.Executor(local_rank)
    ex.ntt_init(max(k - 192, 1), k, n)
    stream = torch.cuda.Stream(device=dev)
    peaks, peak_src = measured_peaks()

    with torch.cuda.stream(stream):
        ex.use_torch_stream()
        witness = torch.empty(R * k * 8, dtype=torch.int32, device=dev)          # [R][k][8] u32
        wbuf = ex.wrap(witness)
        ex.synth(wbuf, SEED, rank * R, R, k)                                     # shard `rank` of the global matrix
        digests = ex.make_device_buffer(n * 32)
        nleaves = world * n
        nodes = ex.make_device_buffer(ex.merkle_node_count(nleaves) * 32)
        gathered = torch.empty(world * n * 8, dtype=torch.int32, device=dev) if world > 1 else None

        exact_engine = None
        if args.layout == "exact" and world > 1:
            spec = importlib.util.spec_from_file_location("lgr_sharding", os.path.join(ROOT, "ligero-prover_b200", "sharding.py"))
            sh = importlib.util.module_from_spec(spec); spec.loader.exec_module(sh)
            T_ex = max(2, (1 << 23) // n)
            exact_engine = sh.GpuEngine(ex, T_ex, world)
            nodes = ex.make_device_buffer(ex.merkle_node_count(n) * 32)

        def step():
            if world == 1:
                ex.encode_commit(wbuf, R, digests, nodes)
            elif exact_engine is not None:
                # global tile i*world + rank = local rows [i*T, (i+1)*T)
                leaves = sh.commit_exact(exact_engine, lambda i: (wbuf.slice(i * T_ex * k * 32, (i + 1) * T_ex * k * 32), T_ex), world * R, T_ex, world, rank, dist)
                ex.use_torch_stream()
                ex.merkle_build(ex.wrap(leaves.contiguous()), n, nodes)
            else:
                ex.encode_commit(wbuf, R, digests, None)
                dist.all_gather_into_tensor(gathered, digests.storage[: n * 8])
                ex.merkle_build(ex.wrap(gathered), nleaves, nodes)

        def sync_all():
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        for _ in range(args.warmup):
            step()
        sync_all()
        ex.profile(True)
        l0 = ex.launch_count()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        sync_all()
        clocks = sampler.stop() if rank == 0 else {}
        ms_total = e0.elapsed_time(e1)
        launches = ex.launch_count() - l0 + (args.steps if world > 1 else 0)
        prof = ex.profile_read()
        ex.profile(False)
        root = ex.copy_to_host(nodes, np.uint8)[:32].tobytes().hex()

        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        ms_step = ms_total / args.steps
        value = world * R * k / (ms_step * 1e-3)

        # ---- end to end: host-resident rows through lgr_encode_commit_host -------------------------
        e2e = None
        if not args.no_e2e:
            need = R * k * 32
            avail = mem_available_bytes()
            e2e_rows = R
            while e2e_rows * k * 32 > max(avail - (8 << 30), 0) * 0.8 // max(world, 1) and e2e_rows > (1 << 12):
                e2e_rows >>= 1
            host = None
            while host is None:
                try:
                    host = torch.empty(e2e_rows * k * 8, dtype=torch.int32, pin_memory=True)
                except RuntimeError:                     # cannot pin that much on this box: halve the e2e witness
                    if e2e_rows <= (1 << 12):
                        raise
                    e2e_rows >>= 1
            host.copy_(witness[: e2e_rows * k * 8])
            stream.synchronize()
            root_e2e = None
            for _ in range(min(args.warmup, 1)):
                _, root_e2e = ex.encode_commit_host(host, e2e_rows)
            sync_all()
            ex.profile(True)
            t0 = time.perf_counter()
            e0.record(stream)
            for _ in range(args.steps):
                _, root_e2e = ex.encode_commit_host(host, e2e_rows)          # blocking: root is on the host on return
            e1.record(stream)
            sync_all()
            wall = (time.perf_counter() - t0) * 1e3
            te = torch.tensor([max(wall, e0.elapsed_time(e1))], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e_ms = float(te.item()) / args.steps
            pe = ex.profile_read()
            ex.profile(False)
            e2e = {"kernel_ms_per_launch": {"encode": pe["encode_ms"] / max(pe["encode_launches"], 1), "sha": pe["sha_ms"] / max(pe["sha_launches"], 1)},"value": world * e2e_rows * k / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": e2e_rows * k * 32,
                   "d2h_bytes_per_step": 32, "ms_per_step": e2e_ms, "rows_per_gpu": e2e_rows,
                   "root_matches_device_path": (root_e2e.hex() == root) if (e2e_rows == R and world == 1) else None}
            del host

        # ---- roofline of the kernel on the critical path + the other kernel of the step ------------
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        kern = {}
        if prof["encode_launches"]:
            per = prof["encode_ms"] / prof["encode_launches"]
            rows_per_launch = R * args.steps / prof["encode_launches"]
            by = rows_per_launch * (k + n) * 32                                  # read k, write n elements per row
            kern["encode_rows_kernel"] = {"ms_per_launch": per, "launches": prof["encode_launches"], "algorithmic_bytes_per_launch": by,
                                          "achieved_gbs": by / (per * 1e-3) / 1e9, "frac_hbm": by / (per * 1e-3) / 1e9 / hbm,
                                          "share_of_step": prof["encode_ms"] / (ms_step * args.steps)}
        if prof["sha_launches"]:
            per = prof["sha_ms"] / prof["sha_launches"]
            rows_per_launch = R * args.steps / prof["sha_launches"]
            by = rows_per_launch * n * 32                                        # every codeword element read once
            split = os.environ.get("LGR_CHAIN_SPLIT")
            lane_split = (n // 16 <= 64) if split is None else (split != "0")       # launch_sha_update's choice (csrc/sha_kernels.cu)
            sha_name = "sha_update_kernel" if n > 18944 else ("sha_chain16_kernel" if (lane_split and n % 16 == 0 and n // 16 <= 148) else "sha_chain_kernel")
            kern[sha_name] = {"ms_per_launch": per, "launches": prof["sha_launches"], "algorithmic_bytes_per_launch": by,
                                         "achieved_gbs": by / (per * 1e-3) / 1e9, "frac_hbm": by / (per * 1e-3) / 1e9 / hbm,
                                         "share_of_step": prof["sha_ms"] / (ms_step * args.steps)}
        dominant = max(kern, key=lambda kk: kern[kk]["share_of_step"]) if kern else None
        roofline = None
        if dominant:
            d = kern[dominant]
            traffic, traffic_src = ncu_traffic(dominant)
            roofline = {"kernel": dominant, "bound": "hbm", "achieved": d["achieved_gbs"], "peak": hbm, "unit": "GB/s", "frac": d["frac_hbm"],
                        "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"],
                        "peak_source": peak_src,
                        "path_bytes_per_element": 32, "path_achieved": value / world * 32 / 1e9, "path_frac": value / world * 32 / 1e9 / hbm,
                        "note": "integer-issue bound path: see int_roofline and DESIGN.md"}
        # the dominant kernel against the bound that actually holds for it: one warp per 32 columns, ALU-issue-bound
        # (12 SHF/LOP3/IADD3 per round at one warp instruction per 2 cycles, 64 rounds: DESIGN.md section 4)
        chain_roofline = None
        chain_name = next((kk for kk in kern if kk.startswith("sha_chain")), None)
        if chain_name and clocks.get("sm_mhz"):
            d = kern[chain_name]
            blocks_per_launch = (R * args.steps / d["launches"]) / 2.0
            cyc = d["ms_per_launch"] * 1e-3 * clocks["sm_mhz"] * 1e6 / blocks_per_launch
            if chain_name == "sha_chain16_kernel":
                floor, why = 1216.0, ("lane-split rounds: the e -> shuffle -> a -> shuffle -> e loop spans 4 rounds and holds two 26-cycle SHFL hops "
                                      "plus two IMAD+IADD3 steps: 19 cycles per round (ALU issue alone would allow 16)")
            else:
                floor, why = 1536.0, "64 rounds x 12 ALU instructions x 2 cycles per warp instruction"
            chain_roofline = {"kernel": chain_name, "bound": "single-warp latency / ALU issue", "unit": "cycles per 64-byte block per column",
                              "floor": floor, "achieved": cyc, "frac": floor / cyc,
                              "note": "a column is one SHA-256 chain: n/32 (n/16) warps of work whatever the GPU; floor = " + why}
        ub = {"imad_wide_per_s": ex.ubench(0), "montmul_per_s": ex.ubench(1), "sha256_compress_per_s": ex.ubench(2)}
        import math
        mm_per_elem = (math.log2(k) - 1) / 2 + 3 + 3 * (math.log2(k) - 1) / 2      # iNTT_k + 3 computed cosets (the 4th is a copy)
        int_roofline = {"measured": ub, "montmul_per_element": mm_per_elem, "sha_compress_per_element": 2,
                        "throughput_bound_elements_per_s": 1.0 / (mm_per_elem / ub["montmul_per_s"] + 2.0 / ub["sha256_compress_per_s"]),
                        "frac": (value / world) / (1.0 / (mm_per_elem / ub["montmul_per_s"] + 2.0 / ub["sha256_compress_per_s"]))}

        aux = {}
        if args.aux and world == 1:
            aux = run_aux(lgr, torch, dev, stream, hbm)

    # ---- CPU baseline: rank 0, N = 1 only; its root pins the GPU path on the same rows ------------
    cpu = None
    parity_check = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, rows, dt = cpu_reference_rate(12.0, k)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "oracle encode+commit of 2^%d rows x k=%d, seed %d, %.1f s, OpenMP threads=%d" % (rows.bit_length() - 1, k, SEED, dt, cores)}
        # the oracle (pinned to the reference's shaders, tests/test_wgslref_cpu.py) committed the first `rows` rows of the
        # bench matrix: commit the same rows on the GPU, device-resident and through the host-buffer entry point
        with torch.cuda.stream(stream):
            ex.use_torch_stream()
            ex.encode_commit(wbuf, rows, digests, nodes)
            gpu_root = ex.copy_to_host(nodes, np.uint8)[:32].tobytes().hex()
            hrows = torch.empty(rows * k * 8, dtype=torch.int32, pin_memory=True)
            hrows.copy_(witness[: rows * k * 8])
            stream.synchronize()
            _, host_root = ex.encode_commit_host(hrows, rows)
        parity_check = {"rows": rows, "k": k, "root_equal": gpu_root == cpu_reference_rate.last_root,
                        "root_equal_host_path": host_root.hex() == cpu_reference_rate.last_root, "oracle_root": cpu_reference_rate.last_root}

    # ---- the bit-exact single-root layout over N GPUs, and the reference's default geometry --------
    exact = exact_k8192 = config5 = k8192 = None
    if not args.no_exact:
        with torch.cuda.stream(stream):
            ex.use_torch_stream()
            if world == 8 and args.log_rows == LOG_ROWS and k == K and not args.no_config5:
                config5 = run_config5(ex, torch, dist, stream, dev, witness, wbuf, rank, world, k, n)
            del wbuf, witness
            torch.cuda.empty_cache()
            if world > 1:
                exact = run_exact(lgr, torch, dist, dev, stream, rank, world, 256, 1 << 20, hbm)
                exact_k8192 = run_exact(lgr, torch, dist, dev, stream, rank, world, 8192, 1 << 15, hbm)
            else:
                k8192 = run_k8192_single(lgr, torch, dev, stream, 1 << 15)
                k8192["per_row_drop_in"] = run_per_row(local_rank)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "ligero encode+commit, R=2^%d rows x k=%d (n=%d) per GPU, BN254 Fr, seed %d" % (args.log_rows, k, n, SEED),
                       "rows_per_gpu": R, "k": k, "n": n, "l2": "inputs (%.0f GiB per GPU) larger than L2 (126 MB); no flush needed" % (R * k * 32 / 2**30),
                       "parallelism": ("rows sharded over %d GPU(s); digests all-gathered (NCCL) for the tree" % world) if exact_engine is None else
                                      ("exact layout over %d GPUs: tiles round-robin, all-to-all of column slabs, one root" % world)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "chain_roofline": chain_roofline, "kernels": kern, "int_roofline": int_roofline,
            "cpu_baseline": cpu, "root": root, "parity_check": parity_check,
        }
        for key, val in (("exact", exact), ("exact_k8192", exact_k8192), ("config5", config5), ("encode_commit_k8192", k8192)):
            if val is not None:
                line[key] = val
        if aux:
            line["aux"] = aux
        print(json.dumps(line), flush=True)
    ex.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def load_sharding():
    spec = importlib.util.spec_from_file_location("lgr_sharding", os.path.join(ROOT, "ligero-prover_b200", "sharding.py"))
    sh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sh)
    return sh


def run_exact(lgr, torch, dist, dev, stream, rank, world, k, total_rows, hbm, steps=3):
    """STRONG scaling of the reference's own commitment: ONE root over a fixed total_rows x k matrix (seed 3), rows dealt
    tile-round-robin over the ranks, column slabs handed to their hashing rank (sharding.commit_exact).  Checked against
    a single-GPU commitment of the same matrix computed on every rank."""
    import numpy as np
    sh = load_sharding()
    n = 4 * k
    ex = lgr.Executor(dev.index)
    ex.ntt_init(max(k - 192, 1), k, n)
    ex.use_torch_stream()
    T = max(2, (1 << 23) // n)
    while T * world > total_rows and T > 2:
        T >>= 1
    num_tiles = (total_rows + T - 1) // T
    mine = sh.tiles_of_rank(num_tiles, world, rank)
    bufs = []
    for t in mine:
        rows = min(T, total_rows - t * T)
        b = ex.make_device_buffer(T * k * 32)
        ex.synth(b, SEED, t * T, rows, k)
        bufs.append((b, rows))
    eng = sh.make_gpu_engine(ex, T, world, rank, dist)
    nodes = ex.make_device_buffer(ex.merkle_node_count(n) * 32)

    def step():
        leaves = sh.commit_exact(eng, lambda i: bufs[i], total_rows, T, world, rank, dist)
        ex.use_torch_stream()
        ex.merkle_build(ex.wrap(leaves.contiguous()), n, nodes)

    def sync_all():
        stream.synchronize()
        dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(2):
        step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    root = ex.copy_to_host(nodes, np.uint8)[:32].tobytes().hex()
    # the same matrix on ONE GPU (every rank does it: no extra collective, and it gives the single-GPU time)
    for b, _ in bufs:
        del b
    bufs.clear()
    whole = ex.make_device_buffer(total_rows * k * 32)
    ex.synth(whole, SEED, 0, total_rows, k)
    d1 = ex.make_device_buffer(n * 32)
    ex.encode_commit(whole, total_rows, d1, nodes)
    stream.synchronize()
    e0.record(stream)
    ex.encode_commit(whole, total_rows, d1, nodes)
    e1.record(stream)
    stream.synchronize()
    single_ms = e0.elapsed_time(e1)
    want = ex.copy_to_host(nodes, np.uint8)[:32].tobytes().hex()
    same = torch.tensor([1 if root == want else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    elems = total_rows * k
    out = {"k": k, "n": n, "rows_total": total_rows, "tile_rows": T, "scaling": "strong", "transport": eng.transport,
           "value": elems / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
           "single_gpu_value": elems / (single_ms * 1e-3), "speedup_vs_single_gpu": single_ms / ms,
           "root_equals_single_gpu": bool(int(same.item())), "root": root,
           "peer_bytes_per_step": int(elems * 4 * 32 * (world - 1) / world),
           "peer_gbs_per_gpu": elems * 4 * 32 * (world - 1) / world / world / (ms * 1e-3) / 1e9,
           "path_frac_hbm": elems / (ms * 1e-3) / world * 32 / 1e9 / hbm}
    eng.close()
    ex.close()
    return out


def run_k8192_single(lgr, torch, dev, stream, total_rows):
    """the reference's default geometry (k = 8192, n = 32768, include/params.hpp:24-32) on one GPU: same fixed matrix as
    exact_k8192 at N > 1 (2^15 rows, seed 3), device-resident and from pinned host rows"""
    import numpy as np
    k, n = 8192, 32768
    ex = lgr.Executor(dev.index)
    ex.ntt_init(k - 192, k, n)
    ex.use_torch_stream()
    w = torch.empty(total_rows * k * 8, dtype=torch.int32, device=dev)
    wb = ex.wrap(w)
    ex.synth(wb, SEED, 0, total_rows, k)
    dig = ex.make_device_buffer(n * 32)
    nodes = ex.make_device_buffer((2 * n - 1) * 32)
    for _ in range(2):
        ex.encode_commit(wb, total_rows, dig, nodes)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream.synchronize()
    e0.record(stream)
    for _ in range(3):
        ex.encode_commit(wb, total_rows, dig, nodes)
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / 3
    root = ex.copy_to_host(nodes, np.uint8)[:32].tobytes().hex()
    host = torch.empty(total_rows * k * 8, dtype=torch.int32, pin_memory=True)
    host.copy_(w)
    stream.synchronize()
    _, r2 = ex.encode_commit_host(host, total_rows)
    t0 = time.perf_counter()
    for _ in range(3):
        _, r2 = ex.encode_commit_host(host, total_rows)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / 3
    elems = total_rows * k
    out = {"k": k, "n": n, "rows_total": total_rows, "value": elems / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "root": root,
           "e2e": {"value": elems / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": elems * 32, "d2h_bytes_per_step": 32,
                   "root_matches_device_path": r2.hex() == root}}
    del host
    ex.close()
    return out


def run_per_row(device_index):
    """the per-row schedule the reference's stage contexts issue (one row per callback, nonbatch_context.hpp:445-451,673-780)
    through the C++ drop-in adapter (ligero::cuda_context): tests/cpp/per_row_bench.cpp, k = 8192"""
    exe = os.path.join(ROOT, "tests", "cpp", "per_row_bench")
    if not os.path.exists(exe):
        return {"unavailable": "tests/cpp/per_row_bench is not built (make -C tests/cpp)"}
    try:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(device_index)))
        out = subprocess.run([exe, "4096", "512"], capture_output=True, text=True, timeout=120, env=env)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:                                   # noqa: BLE001 -- an auxiliary measurement must not sink the bench line
        return {"unavailable": "per_row_bench failed: %s" % e}


def run_config5(ex, torch, dist, stream, dev, witness, wbuf, rank, world, k, n):
    """BASELINE config 5 at its stated size: 2^26 rows x k = 256 over 8 GPUs = 2^23 rows (64 GiB) per GPU, sharded layout
    (per-rank commitments, one digest all-gather, tree over the 8n leaves).  The first 2^22 rows of the shard are the
    resident bench witness; the second half is generated next to it."""
    import numpy as np
    R = 1 << 22
    second = torch.empty(R * k * 8, dtype=torch.int32, device=dev)
    sbuf = ex.wrap(second)
    ex.synth(sbuf, SEED, (world + rank) * R, R, k)           # rows [(8+rank)*2^22, ...): disjoint from every first half
    ctx = ex.make_device_buffer(ex.sha256_context_bytes(n))
    dig = ex.make_device_buffer(n * 32)
    bind = ex.bind_sha256_context(ctx, dig)
    gathered = torch.empty(world * n * 8, dtype=torch.int32, device=dev)
    nodes = ex.make_device_buffer(ex.merkle_node_count(world * n) * 32)
    ex.sha256_init(n)

    def step():
        ex.sha256_digest_init(bind)
        ex.encode_absorb(ctx, wbuf, R)
        ex.encode_absorb(ctx, sbuf, R)
        ex.sha256_digest_final(bind)
        dist.all_gather_into_tensor(gathered, dig.storage[: n * 8])
        ex.merkle_build(ex.wrap(gathered), world * n, nodes)

    step()
    stream.synchronize(); dist.barrier(); torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(2):
        step()
    e1.record(stream)
    stream.synchronize(); dist.barrier(); torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / 2
    root = ex.copy_to_host(nodes, np.uint8)[:32].tobytes().hex()
    del second
    return {"rows_total": world * 2 * R, "rows_per_gpu": 2 * R, "k": k, "layout": "sharded (one digest all-gather)", "value": world * 2 * R * k / (ms * 1e-3),
            "unit": UNIT, "ms_per_step": ms, "steps": 2, "root": root}


def run_aux(lgr, torch, dev, stream, hbm):
    """extra measurements (not the headline): reference default geometry and BASELINE config 2"""
    out = {}
    # reference default geometry k = 8192, n = 32768 (include/params.hpp:24-32), 2^14 rows = 4 GiB
    k, R = 8192, 1 << 14
    ex = lgr.Executor(dev.index)
    ex.ntt_init(k - 192, k, 4 * k)
    ex.use_torch_stream()
    w = torch.empty(R * k * 8, dtype=torch.int32, device=dev)
    wb = ex.wrap(w)
    ex.synth(wb, 5, 0, R, k)
    dig = ex.make_device_buffer(4 * k * 32); nodes = ex.make_device_buffer((8 * k - 1) * 32)
    for _ in range(2):
        ex.encode_commit(wb, R, dig, nodes)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream.synchronize(); e0.record(stream)
    for _ in range(3):
        ex.encode_commit(wb, R, dig, nodes)
    e1.record(stream); stream.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out["encode_commit_k8192"] = {"rows": R, "k": k, "ms_per_step": ms, "elements_per_s": R * k / (ms * 1e-3)}
    # BASELINE config 2: 2^20-point NTT then iNTT, 64 MiB algorithmic bytes per transform
    logn = 20
    buf = ex.make_device_buffer((1 << logn) * 32)
    ex.synth(buf, 2, 0, 1, 1 << logn)
    wroot = lgr.root_of_unity(logn)
    for _ in range(3):
        ex.ntt_pow2(buf, logn, 1, wroot, False); ex.ntt_pow2(buf, logn, 1, wroot, True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tf = ti = 0.0
    for _ in range(5):
        flush.fill_(1); e0.record(stream); ex.ntt_pow2(buf, logn, 1, wroot, False); e1.record(stream); stream.synchronize(); tf += e0.elapsed_time(e1)
        flush.fill_(2); e0.record(stream); ex.ntt_pow2(buf, logn, 1, wroot, True); e1.record(stream); stream.synchronize(); ti += e0.elapsed_time(e1)
    by = 2 * (1 << logn) * 32
    for name, tms in (("ntt_2^20_forward", tf / 5), ("ntt_2^20_inverse", ti / 5)):
        out[name] = {"ms": tms, "algorithmic_bytes": by, "achieved_gbs": by / (tms * 1e-3) / 1e9, "frac_hbm": by / (tms * 1e-3) / 1e9 / hbm, "l2": "flushed (256 MiB write) before each timed launch"}
    ex.close()
    out["prove_k256"] = run_prove_aux(lgr, dev)
    return out


def run_prove_aux(lgr, dev, k=256, l=64, n_linear=1 << 14, n_quad=1 << 13):
    """the three-stage matrix prover (include/lgr_prover.h) on a synthetic satisfiable statement: witness
    elements per second through all three stages, host-resident rows in, gzip proof out"""
    import importlib
    import numpy as np
    pr = importlib.import_module("ligero_prover_b200.prover")
    rng = np.random.default_rng(4)
    kinds = np.concatenate([np.zeros(n_linear, np.uint8), np.ones(n_quad, np.uint8)])
    rng.shuffle(kinds)
    rows = int(kinds.size + 2 * kinds.sum())
    values = np.zeros((rows, l, 8), np.uint32)
    values[:, :, :2] = rng.integers(0, 1 << 32, size=(rows, l, 2), dtype=np.uint32)          # 64-bit witnesses
    starts = np.cumsum(np.where(kinds == 1, 3, 1)) - np.where(kinds == 1, 3, 1)
    qx = starts[kinds == 1]
    x = values[qx, :, 0].astype(np.uint64) | (values[qx, :, 1].astype(np.uint64) << np.uint64(32))
    y = values[qx + 1, :, 0].astype(np.uint64) | (values[qx + 1, :, 1].astype(np.uint64) << np.uint64(32))
    # z = x*y < 2^128 < p: exact 128-bit products from 32-bit halves
    xl, xh, yl, yh = (x & np.uint64(0xFFFFFFFF)), (x >> np.uint64(32)), (y & np.uint64(0xFFFFFFFF)), (y >> np.uint64(32))
    ll, lh, hl, hh = xl * yl, xl * yh, xh * yl, xh * yh
    mid = (ll >> np.uint64(32)) + (lh & np.uint64(0xFFFFFFFF)) + (hl & np.uint64(0xFFFFFFFF))
    z = np.zeros((qx.size, l, 8), np.uint32)
    z[:, :, 0] = (ll & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    z[:, :, 1] = (mid & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = hh + (lh >> np.uint64(32)) + (hl >> np.uint64(32)) + (mid >> np.uint64(32))
    z[:, :, 2] = (hi & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    z[:, :, 3] = (hi >> np.uint64(32)).astype(np.uint32)
    values[qx + 2] = z
    ex = lgr.Executor(dev.index)
    ex.ntt_init(l, k, 4 * k)
    ex.use_torch_stream()
    proof = pr.prove(ex, kinds, values, None, 0, bytes(range(32)))            # warm-up (tables, allocations)
    proof.close()
    t0 = time.perf_counter()
    proof = pr.prove(ex, kinds, values, None, 0, bytes(range(32)))
    dt = time.perf_counter() - t0
    info, tm = proof.info(), proof.timing()
    res = {"k": k, "l": l, "linear_rows": int(n_linear), "quadratic_triples": int(n_quad), "encoded_rows": int(info["encoded_rows"]),
           "valid": list(info["valid"]), "seconds": dt, "witness_elements_per_s": rows * l / dt, "padded_elements_per_s": rows * k / dt,
           "proof_bytes_gzip": len(proof.gzip), **tm}
    proof.close()
    ex.close()
    return res


if __name__ == "__main__":
    sys.exit(main())
