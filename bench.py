#!/usr/bin/env python
"""Benchmark of the VB E-step on BASELINE.json's headline config (configs[2]):
synthetic D=1M docs, V=100k, K=100, Zipf document lengths.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--docs D] [--state cold|warm]

A "step" is one full E-step (variational_bayes.py:132-216) over the rank's corpus: the
E_log_eta producer, the per-document kernels, the ELBO reduction and -- with N > 1 -- the NCCL
all-reduce of the K x V statistics.  Weak scaling: every rank holds D documents (different
seeds), so N GPUs process N*D documents per step.

value : docs/s with corpus, eta and alpha already resident in HBM (pylda_estep_resident)
e2e   : docs/s through the reference-facing call pylda_estep with HOST buffers: H2D of
        eta/alpha and D2H of gamma, phi_ss and the ELBO inside the timed region
state : "cold" = eta0 ~ Gamma(100, 1/100) (EM iteration 1, nearly every document runs to the
        50-trip cap -- the worst case); the JSON also carries the same measurement at a warm
        state (EM iteration 5: four resident EM iterations incl. the alpha update) under "warm".

--impl reference times the CPU restatement of the reference (oracle/estep_oracle.py, same numpy
call sequence as the reference) on all host cores over a bounded sample of the same corpus.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy  # noqa: E402

K_TOPICS, V_TYPES, D_DOCS = 100, 100000, 1000000
METRIC = "E-step docs/sec at K=100, V=100k (synthetic Zipf-length corpus)"
UNIT = "docs/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_corpus(D, V, seed):
    """Seeded synthetic corpus, cached under /tmp so the N=1,2,4,8 runs on one box reuse it."""
    from pylda_b200 import synthetic
    path = "/tmp/pylda_bench_D%d_V%d_s%d.npz" % (D, V, seed)
    if os.path.exists(path):
        try:
            z = numpy.load(path)
            return z["row_ptr"], z["ids"], z["cts"]
        except Exception:
            pass
    t = time.time()
    row_ptr, ids, cts = synthetic.synthetic_corpus(D, V, seed=seed, length="zipf")
    log("generated corpus D=%d nnz=%d in %.1fs" % (D, len(ids), time.time() - t))
    tmp = path + ".%d.tmp.npz" % os.getpid()
    numpy.savez(tmp, row_ptr=row_ptr, ids=ids, cts=cts)
    os.replace(tmp, path)
    return row_ptr, ids, cts


class ClockSampler(object):
    """nvidia-smi clock / throttle-reason sampler for the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(numpy.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_traffic(D):
    """dram__bytes_read.sum + dram__bytes_write.sum of the per-document kernels of one E-step, from the
    committed ncu --set full capture of this very workload (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_per_estep"]) if int(t["docs"]) == int(D) else None
    except Exception:
        return None


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
# reference arm: CPU restatement on all host cores
# ------------------------------------------------------------------------------------------
_W = {}


def _worker_estep(args):
    lo, hi = args
    from oracle import estep_oracle as O
    rp = _W["row_ptr"][lo:hi + 1] - _W["row_ptr"][lo]
    a, b = int(_W["row_ptr"][lo]), int(_W["row_ptr"][hi])
    t = time.perf_counter()
    r = O.e_step(rp, _W["ids"][a:b], _W["cts"][a:b], _W["eta"], _W["alpha"], 50, 1e-6, return_iters=True)
    return hi - lo, time.perf_counter() - t, int(r["iters"].sum()), r["doc_ll"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    import multiprocessing as mp
    from pylda_b200 import synthetic
    cores = os.cpu_count() or 1
    per_worker = args.ref_docs_per_core
    sample_docs = cores * per_worker
    D = max(sample_docs * (args.warmup + args.steps), 4096)
    D = min(D, args.docs)
    row_ptr, ids, cts = load_corpus(args.docs, V_TYPES, 1236)
    _W.update(row_ptr=row_ptr, ids=ids, cts=cts, eta=synthetic.initial_eta(K_TOPICS, V_TYPES, 0),
              alpha=numpy.full(K_TOPICS, 1.0 / K_TOPICS))
    ctxm = mp.get_context("fork")
    times, docs, iters = [], [], []
    with ctxm.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            base = (step * sample_docs) % max(1, D - sample_docs + 1)
            jobs = [(base + w * per_worker, base + (w + 1) * per_worker) for w in range(cores)]
            t = time.perf_counter()
            res = pool.map(_worker_estep, jobs)
            dt = time.perf_counter() - t
            if step >= args.warmup:
                times.append(dt)
                docs.append(sum(r[0] for r in res))
                iters.append(sum(r[2] for r in res))
    total_t, total_d = sum(times), sum(docs)
    value = total_d / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[2]: synthetic D=%d docs per GPU, V=%d, K=%d, Zipf lengths (nnz=%d on rank 0)" % (
            args.docs, V_TYPES, K_TOPICS, len(ids)),
            "state": "cold (eta0 ~ Gamma(100,0.01), EM iteration 1)",
            "mean_inner_trips": sum(iters) / max(1, total_d),
            "local_parameter_iteration": 50, "converge_threshold": 1e-06,
            "parallelism": "%d host processes, one document shard each (the reference itself is single-threaded)" % cores,
            "timing": "wall clock around each bounded sample step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d docs/step (%d per core) of the same corpus, %d worker processes of the numpy "
                                   "restatement of variational_bayes.py:132-216 (the Python-2 reference cannot run "
                                   "on the box); mean inner trips %.1f" % (
                                       sample_docs, per_worker, cores, sum(iters) / max(1, total_d))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))
    return 0


def cpu_baseline_single_core(row_ptr, ids, cts, eta, alpha, ndocs=2000):
    from oracle import estep_oracle as O   # checker used as the reported CPU baseline only
    rp = row_ptr[:ndocs + 1]
    nz = int(rp[-1])
    t = time.perf_counter()
    r = O.e_step(rp, ids[:nz], cts[:nz], eta, alpha, 50, 1e-6, return_iters=True)
    dt = time.perf_counter() - t
    return {"value": ndocs / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "first %d docs of the same corpus, eta0, 1 process / 1 thread (the reference is single-threaded); "
                      "mean inner trips %.1f; %.1f s" % (ndocs, float(r["iters"].mean()), dt)}


# ------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------
def run_product(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist   # host-side rendezvous only (gloo); the data path is NCCL inside the library
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)

    from pylda_b200 import native, synthetic
    ctx = native.EStepContext(local_rank)
    if world > 1:
        ids_obj = [native.EStepContext.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids_obj, src=0)
        ctx.comm_init(world, rank, ids_obj[0])

    K, V, D = K_TOPICS, V_TYPES, args.docs
    row_ptr, ids, cts = load_corpus(D, V, 1236 + rank)
    eta0 = synthetic.initial_eta(K, V, 0)
    alpha = numpy.full(K, 1.0 / K)
    alpha_beta = 1.0 / V
    ctx.set_corpus(0, row_ptr, ids, cts)
    nnz = int(len(ids))

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    def timed_resident(nsteps):
        """K steps with everything resident; returns (wall_s, device_ms list, kernel_ms list, last stats)."""
        barrier()
        t0 = time.perf_counter()
        dev, ker, st = [], [], None
        for _ in range(nsteps):
            st = ctx.estep_resident(0, 50, 1e-6)
            dev.append(st["total_ms"]); ker.append(st["kernel_ms"])
        wall = time.perf_counter() - t0
        barrier()
        return wall, dev, ker, st

    def warm_model(want_eta):
        """EM iterations 1..4 of variational_bayes.py:239-261 with everything resident: device E-step,
        device M-step, alpha statistics from the device (summed over ranks by the library) and the
        reference's Newton update of alpha on the host (K numbers).  Leaves the model of EM iteration 5
        on the device; returns (eta or None, alpha)."""
        from pylda_b200.variational_bayes import VariationalBayes
        shell = VariationalBayes()
        shell._number_of_topics = K
        shell._number_of_documents = int(allsum(float(D)))
        shell._alpha_alpha = alpha.copy()
        ctx.set_model(eta0, alpha)
        eta_host = None
        em_wall = []
        for em in range(4):
            t_em = time.perf_counter()
            ctx.estep_resident(0, 50, 1e-6, want_alpha_ss=True)
            alpha_ss = ctx.get_results(0, gamma=False, phi=False, alpha_ss=True)["alpha_ss"]
            _, eta_host = ctx.mstep_resident(alpha_beta, want_eta=(want_eta and em == 3))
            shell.optimize_hyperparameters(alpha_ss)
            ctx.set_alpha(shell._alpha_alpha)
            em_wall.append(time.perf_counter() - t_em)
        em_stats["ms"] = 1e3 * min(em_wall[:3])      # the 4th may include the eta copy-back
        return eta_host, shell._alpha_alpha.copy()

    em_stats = {}
    results = {}
    sampler = ClockSampler(local_rank)
    alpha_warm = alpha
    for state in ("cold", "warm"):
        if state == "warm":
            _, alpha_warm = warm_model(False)
        else:
            ctx.set_model(eta0, alpha)
        for _ in range(args.warmup):
            ctx.estep_resident(0, 50, 1e-6)
        if state == args.state:
            sampler.start()
        wall, dev, ker, st = timed_resident(args.steps)
        if state == args.state:
            clocks = sampler.stop()
        dev_ms = allmax(sum(dev) / len(dev))
        wall_ms = allmax(1e3 * wall / args.steps)
        ker_ms = sum(ker) / len(ker)
        docs_total = allsum(float(D))
        res = ctx.get_results(0, gamma=False, phi=False)
        results[state] = dict(dev_ms=dev_ms, wall_ms=wall_ms, ker_ms=ker_ms, docs_total=docs_total, stats=st,
                              doc_ll=res["doc_ll"], mean_trips=st["inner_iters"] / docs_total,
                              at_cap=st["docs_at_cap"])

    # ---- memory-path probe: the same E-step limited to ONE trip per document (local_parameter_iteration=1):
    # gather + one fixed-point trip + scatter, i.e. the regime where the HBM roof binds ----
    ctx.set_model(eta0, alpha)
    for _ in range(2):
        ctx.estep_resident(0, 1, 1e-6)
    probe_ms = sorted(ctx.estep_resident(0, 1, 1e-6)["kernel_ms"] for _ in range(3))[1]
    probe_ms = allmax(probe_ms)

    # ---- e2e: the reference-facing call with host buffers (pinned), at the headline state ----
    alpha_e2e = alpha
    if args.state == "warm":
        eta_host, alpha_e2e = warm_model(True)
    else:
        eta_host = eta0
    eta_pin = numpy.ascontiguousarray(eta_host)
    gamma_pin = numpy.empty((D, K), dtype=numpy.float64)
    phi_pin = numpy.empty((K, V), dtype=numpy.float64)
    gamma_pin.fill(0.0); phi_pin.fill(0.0)
    for a in (eta_pin, gamma_pin, phi_pin):
        ctx.pin(a)
    e2e_steps = max(2, min(args.steps, 5))
    ctx.estep(0, eta_pin, alpha_e2e, 50, 1e-6, gamma_out=gamma_pin, phi_out=phi_pin)   # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = ctx.estep(0, eta_pin, alpha_e2e, 50, 1e-6, gamma_out=gamma_pin, phi_out=phi_pin)
    e2e_wall = time.perf_counter() - t0
    barrier()
    e2e_ms = allmax(1e3 * e2e_wall / e2e_steps)
    h2d = eta_pin.nbytes + alpha.nbytes
    d2h = gamma_pin.nbytes + phi_pin.nbytes + 64
    e2e_doc_ll = out["doc_ll"]
    for a in (eta_pin, gamma_pin, phi_pin):
        ctx.unpin(a)

    head = results[args.state]
    st = head["stats"]
    peak, peak_src = measured_peak_hbm()
    algo = st["algo_total_bytes"]
    achieved = algo / (head["ker_ms"] * 1e-3) / 1e9
    achieved_read = st["algo_read_bytes"] / (head["ker_ms"] * 1e-3) / 1e9

    # secondary roofline: the fp64 pipe.  Algorithmic flops = 4*K per (row, trip) for the two mat-vecs
    # (exp(psi), reciprocals and reductions not counted); peak = DFMA rate measured on this part by
    # scripts/ubench/fp64_lat.cu (58.8 lanes/clk/SM * 148 SMs * 1.965 GHz * 2 flop = 34.2 TFLOP/s)
    fp64_peak = 58.8 * 148 * 1.965e9 * 2 / 1e12
    fp64_flops = 4.0 * K * st["row_trips"] / max(1, world)      # row_trips is summed over ranks
    fp64_ach = fp64_flops / (head["ker_ms"] * 1e-3) / 1e12

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_single_core(row_ptr, ids, cts, eta_host, alpha_e2e, args.cpu_docs)

    if rank == 0:
        other = "warm" if args.state == "cold" else "cold"
        o = results[other]
        line = {
            "metric": METRIC, "value": head["docs_total"] / (head["dev_ms"] * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["dev_ms"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "configs[2]: synthetic D=%d docs per GPU, V=%d, K=%d, Zipf lengths (nnz=%d on rank 0)" % (
                    D, V, K, nnz),
                "state": args.state + (" (eta0 ~ Gamma(100,0.01), EM iteration 1)" if args.state == "cold"
                                       else " (EM iteration 5: after 4 EM iterations with alpha updates)"),
                "mean_inner_trips": head["mean_trips"], "docs_at_cap": head["at_cap"],
                "local_parameter_iteration": 50, "converge_threshold": 1e-6,
                "l2": "inputs larger than L2 (CSR + gamma + tables = %.1f GB per GPU)" % (
                    (12.0 * nnz + 8.0 * D * (K + 2) + 4 * 8.0 * V * K) / 1e9),
                "parallelism": "dp%d, one process per GPU, one NCCL all-reduce of K x V f64 per step" % world,
                "timing": "CUDA events on the library's stream, per step, max over ranks",
            },
            "wall_ms_per_step": head["wall_ms"],
            "elbo_doc_ll": head["doc_ll"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(D), "peak_source": peak_src,
                         "kernel": "per-document E-step kernels of one E-step (estep_rt / estep_v2 / streaming, "
                                   "one launch per length class)",
                         "kernel_ms": head["ker_ms"], "achieved_read": achieved_read,
                         "frac_read": achieved_read / peak,
                         "note": "binding roof at %.1f trips/doc is the fp64 pipe, not HBM (DESIGN.md 4.1); "
                                 "see roofline_fp64" % head["mean_trips"]},
            "roofline_fp64": {"bound": "fp64 pipe", "achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": fp64_ach / fp64_peak,
                              "flops": "4*K per (term row, trip) of the two mat-vecs; exp(psi) and reductions not counted",
                              "peak_source": "measured DFMA rate, scripts/ubench/fp64_lat.cu (profiles/r1_fp64_ubench.txt)"},
            other: {"value": o["docs_total"] / (o["dev_ms"] * 1e-3), "ms_per_step": o["dev_ms"], "kernel_ms": o["ker_ms"],
                    "mean_inner_trips": o["mean_trips"], "elbo_doc_ll": o["doc_ll"],
                    "roofline_frac": o["stats"]["algo_total_bytes"] / (o["ker_ms"] * 1e-3) / 1e9 / peak,
                    "roofline_frac_read": o["stats"]["algo_read_bytes"] / (o["ker_ms"] * 1e-3) / 1e9 / peak},
            "em_iteration": {"what": "one whole resident EM iteration as VariationalBayes.learning() runs it: E-step + alpha "
                                     "statistics + device M-step + host Newton update of alpha (wall clock, rank 0)",
                             "ms": em_stats.get("ms")},
            "probe_1trip": {"what": "same corpus with local_parameter_iteration=1 (gather + one trip + scatter): the "
                                    "regime where the HBM roof binds", "kernel_ms": probe_ms,
                            "achieved": algo / (probe_ms * 1e-3) / 1e9, "unit": "GB/s",
                            "frac": algo / (probe_ms * 1e-3) / 1e9 / peak},
            "e2e": {"value": head["docs_total"] / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "elbo_doc_ll": e2e_doc_ll,
                    "host_buffers": "page-locked (cudaHostRegister, mapped): eta H2D and phi_ss D2H by cudaMemcpyAsync, "
                                    "gamma stored straight into the host buffer by the kernels (counted in d2h bytes)"},
            "gpu_launches": int(st["n_launches"]) * args.steps,
            "estep_kernel_launches_per_step": int(st["n_estep_launches"]),
            "docs_resident": st["docs_resident"], "docs_streamed": st["docs_streamed"],
            "clocks": clocks,
            "device": ctx.device_name(),
        }
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        emit(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


class StdoutGuard(object):
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its
    version banner to stdout at communicator creation), so fd 1 is pointed at stderr for the whole
    run and the JSON line goes to the saved, real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


GUARD = None


def emit(line):
    if GUARD is not None:
        GUARD.emit(line)
    else:
        print(line, flush=True)


def main():
    global GUARD
    GUARD = StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--docs", type=int, default=D_DOCS, help="documents per GPU")
    ap.add_argument("--state", default="cold", choices=["cold", "warm"])
    ap.add_argument("--cpu-docs", type=int, default=2000)
    ap.add_argument("--ref-docs-per-core", type=int, default=150)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("note: W >= 3 warm-up steps are required for a valid number; got %d" % args.warmup)
    if args.impl == "reference":
        return run_reference(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
