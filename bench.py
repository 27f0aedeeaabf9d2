#!/usr/bin/env python
"""Benchmark of the VB E-step (variational_bayes.py:132-216) on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c3|c5|c5x]
                    [--scaling weak|strong] [--docs D] [--state cold|warm]

Workloads (`--config`):
  c2   configs[1]: synthetic D = 100k docs, V = 10k, K = 50, ~100 tokens per document
  c3   BASELINE.json configs[2], the headline: synthetic D = 1M docs, V = 100k, K = 100, Zipf document lengths
  c5   configs[4]: D = 1.25M docs per GPU, V = 1M, K = 500, ~100 tokens per document
  c5x  configs[4] contention stress: the same with word exponent 1.3 (hot words hit by most documents)

A "step" is one full E-step over the rank's corpus: the E_log_eta producer, the per-document kernels, the
ELBO reduction and -- with N > 1 -- the NCCL all-reduce of the K x V statistics.
  --scaling weak   (default) every rank holds D documents (different seeds): N GPUs process N*D per step
  --scaling strong the ONE corpus of D documents is cut into N nnz-balanced contiguous shards

value : docs/s with corpus, eta and alpha already resident in HBM (pylda_estep_resident)
e2e   : docs/s through the reference-facing call pylda_estep with HOST buffers: H2D of eta/alpha and D2H of
        gamma, phi_ss and the ELBO inside the timed region
state : "cold" = eta0 ~ Gamma(100, 1/100) (EM iteration 1, nearly every document runs to the 50-trip cap -- the
        worst case); "warm" = EM iteration 5 of the same corpus; the line also carries `warm_lda`: EM iteration 5
        on a corpus drawn from an actual LDA model, where the fixed point stops after ~10 trips.
check : every run verifies, on the results of the timed state, that sum(phi_ss) == number of tokens over all ranks
        and sum_k gamma_dk == sum(alpha) + N_d for every document (residuals in the line).

--impl reference times the reference's own CPU path on all host cores over a bounded sample of the same corpus:
the UNMODIFIED reference (parse_data + e_step through oracle/ref_shim.py) when its sources are present
($PYLDA_REF, baseline/_ref, /root/reference), else the numpy restatement oracle/estep_oracle.py.
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

METRIC_C3 = "E-step docs/sec at K=100, V=100k (synthetic Zipf-length corpus)"
UNIT = "docs/s"

CONFIGS = {
    "c2": dict(K=50, V=10000, D=100000, length="poisson", mean_len=100, word_exponent=1.0, seed=1235,
               metric="E-step docs/sec at K=50, V=10k (synthetic ~100-token documents)",
               label="configs[1]: synthetic D=%(D)d docs%(per)s, V=%(V)d, K=%(K)d, ~100 tokens per document (nnz=%(nnz)d on rank 0)"),
    "c3": dict(K=100, V=100000, D=1000000, length="zipf", mean_len=100, word_exponent=1.0, seed=1236, metric=METRIC_C3,
               label="configs[2]: synthetic D=%(D)d docs%(per)s, V=%(V)d, K=%(K)d, Zipf lengths (nnz=%(nnz)d on rank 0)"),
    "c5": dict(K=500, V=1000000, D=1250000, length="poisson", mean_len=100, word_exponent=1.0, seed=1238,
               metric="E-step docs/sec at K=500, V=1M (synthetic ~100-token documents)",
               label="configs[4]: synthetic D=%(D)d docs%(per)s, V=%(V)d, K=%(K)d, ~100 tokens per document (nnz=%(nnz)d on rank 0)"),
    "c5x": dict(K=500, V=1000000, D=1250000, length="poisson", mean_len=100, word_exponent=1.3, seed=1238,
                metric="E-step docs/sec at K=500, V=1M (synthetic ~100-token documents, word exponent 1.3)",
                label="configs[4] contention stress: synthetic D=%(D)d docs%(per)s, V=%(V)d, K=%(K)d, ~100 tokens per document, "
                      "word exponent 1.3 (nnz=%(nnz)d on rank 0)"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_corpus(D, V, seed, length="zipf", mean_len=100, word_exponent=1.0, kind="zipf1"):
    """Seeded synthetic corpus, cached under /tmp so the N=1,2,4,8 runs on one box reuse it."""
    from pylda_b200 import synthetic
    tag = "lda" if kind == "lda" else "%s_m%d_x%g" % (length, mean_len, word_exponent)
    path = "/tmp/pylda_bench_D%d_V%d_s%d_%s.npz" % (D, V, seed, tag)
    if os.path.exists(path):
        try:
            z = numpy.load(path)
            return z["row_ptr"], z["ids"], z["cts"]
        except Exception:
            pass
    t = time.time()
    if kind == "lda":
        row_ptr, ids, cts = synthetic.lda_corpus(D, V, seed)
    else:
        row_ptr, ids, cts = synthetic.synthetic_corpus(D, V, seed=seed, length=length, mean_len=mean_len,
                                                       word_exponent=word_exponent)
    log("generated corpus D=%d nnz=%d in %.1fs" % (D, len(ids), time.time() - t))
    tmp = path + ".%d.tmp.npz" % os.getpid()
    numpy.savez(tmp, row_ptr=row_ptr, ids=ids, cts=cts)
    os.replace(tmp, path)
    return row_ptr, ids, cts


def rank_corpus(cfg, D, rank, world, scaling):
    """The documents of this rank.  weak: its own corpus of D documents (seed + rank); strong: its nnz-balanced
    contiguous shard of the one corpus of D documents."""
    from pylda_b200 import native
    kw = dict(length=cfg["length"], mean_len=cfg["mean_len"], word_exponent=cfg["word_exponent"])
    if scaling == "weak" or world == 1:
        return load_corpus(D, cfg["V"], cfg["seed"] + rank, **kw)
    row_ptr, ids, cts = load_corpus(D, cfg["V"], cfg["seed"], **kw)
    b = native.shard_bounds(row_ptr, world)
    return native.shard_csr(row_ptr, ids, cts, int(b[rank]), int(b[rank + 1]))


def shared_config(args, cfg, nnz0):
    """The `config` object: identical keys and values in the product arm and the reference arm."""
    per = " per GPU" if args.scaling == "weak" else " in total"
    return {
        "workload": cfg["label"] % dict(D=args.docs, V=cfg["V"], K=cfg["K"], nnz=nnz0, per=per),
        "state": args.state + (" (eta0 ~ Gamma(100,0.01), EM iteration 1)" if args.state == "cold"
                               else " (EM iteration 5: after 4 EM iterations with alpha updates)"),
        "local_parameter_iteration": 50, "converge_threshold": 1e-6,
        "scaling": args.scaling,
    }


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


def build_digest():
    try:
        with open(os.path.join(ROOT, "pylda_b200", "csrc", "build", "stamp")) as f:
            return f.read().strip()
    except Exception:
        return None


def measured_traffic(config_name, D):
    """dram__bytes_read.sum + dram__bytes_write.sum of the per-document kernels of one E-step, from the committed
    ncu --set full capture of this very workload (profiles/traffic.json, written by scripts/ncu_traffic.py), or
    None.  Returns (bytes, note): the note says whether the capture was taken from the build that is running."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if int(t["docs"]) != int(D) or t.get("config", "c3") != config_name:
            return None, "no capture for this workload"
        same = t.get("build_digest") == build_digest()
        return float(t["dram_bytes_per_estep"]), ("ncu capture of this build" if same else
                                                    "ncu capture of an earlier build of the same kernels (digest differs)")
    except Exception:
        return None, "profiles/traffic.json missing"


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU path on all host cores
# ------------------------------------------------------------------------------------------
_W = {}


def _worker_port(job):
    lo, hi = job
    from oracle import estep_oracle as O
    rp = _W["row_ptr"][lo:hi + 1] - _W["row_ptr"][lo]
    a, b = int(_W["row_ptr"][lo]), int(_W["row_ptr"][hi])
    t = time.perf_counter()
    r = O.e_step(rp, _W["ids"][a:b], _W["cts"][a:b], _W["eta"], _W["alpha"], 50, 1e-6, return_iters=True)
    return hi - lo, time.perf_counter() - t, int(r["iters"].sum())


def _worker_shim(job):
    """The unmodified reference on a shard rendered as text: parse_data + _initialize outside the clock (they are
    not on the path), VariationalBayes.e_step (variational_bayes.py:132-216) inside."""
    lo, hi = job
    import contextlib
    import io
    from oracle import ref_shim
    from pylda_b200 import synthetic
    _, ref_vb = ref_shim.load()
    rp = _W["row_ptr"][lo:hi + 1] - _W["row_ptr"][lo]
    a, b = int(_W["row_ptr"][lo]), int(_W["row_ptr"][hi])
    docs = synthetic.render_text(rp, _W["ids"][a:b], _W["cts"][a:b])
    K, V = _W["eta"].shape
    with contextlib.redirect_stdout(io.StringIO()):
        vb = ref_vb.VariationalBayes()
        numpy.random.seed(0)
        vb._initialize(docs, _W["vocab"], K, 1.0 / K, 1.0 / V)
        count = [0]
        real_tile = numpy.tile

        def counting_tile(*x, **k):          # one numpy.tile call per inner trip (:177)
            count[0] += 1
            return real_tile(*x, **k)
        numpy.tile = counting_tile
        try:
            t = time.perf_counter()
            vb.e_step()
            dt = time.perf_counter() - t
        finally:
            numpy.tile = real_tile
    return hi - lo, dt, count[0]


def reference_runner(cfg, args):
    """(worker function, kind, description) of the CPU arm: the real reference when its sources are present."""
    from oracle import ref_shim
    if ref_shim.available() and not args.ref_port and cfg["K"] * cfg["V"] <= 5e7:    # (K x V draw per worker: c3 only)
        return _worker_shim, "reference", ("the UNMODIFIED reference (%s: parse_data + e_step, loaded through oracle/ref_shim.py)"
                                           % ref_shim._find_root())
    return _worker_port, "port", ("numpy restatement of variational_bayes.py:132-216 (oracle/estep_oracle.py; the reference "
                                  "sources are not on this box)")


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    import multiprocessing as mp
    from pylda_b200 import synthetic
    K, V = cfg["K"], cfg["V"]
    cores = os.cpu_count() or 1
    per_worker = args.ref_docs_per_core
    sample_docs = cores * per_worker
    row_ptr, ids, cts = load_corpus(args.docs, V, cfg["seed"], cfg["length"], cfg["mean_len"], cfg["word_exponent"])
    D = len(row_ptr) - 1
    worker, kind, what = reference_runner(cfg, args)
    _W.update(row_ptr=row_ptr, ids=ids, cts=cts, eta=synthetic.initial_eta(K, V, 0), alpha=numpy.full(K, 1.0 / K),
              vocab=["w%d" % i for i in range(V)])
    ctxm = mp.get_context("fork")
    times, docs, iters = [], [], []
    with ctxm.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            base = (step * sample_docs) % max(1, D - sample_docs + 1)
            jobs = [(base + w * per_worker, base + (w + 1) * per_worker) for w in range(cores)]
            t = time.perf_counter()
            res = pool.map(worker, jobs)
            dt = time.perf_counter() - t
            if kind == "reference":
                dt = max(r[1] for r in res)          # the e_step clocks of the workers (parsing is not on the path)
            if step >= args.warmup:
                times.append(dt)
                docs.append(sum(r[0] for r in res))
                iters.append(sum(r[2] for r in res))
    total_t, total_d = sum(times), sum(docs)
    value = total_d / total_t
    trips = sum(iters) / max(1, total_d)
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(args, cfg, len(ids)),
        "run": {"mean_inner_trips": trips,
                "parallelism": "%d host processes, one document shard each (the reference itself is single-threaded)" % cores,
                "timing": "wall clock around each bounded sample step" if kind == "port"
                          else "slowest worker's clock around VariationalBayes.e_step per bounded sample step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d docs/step (%d per core) of the same corpus, %d worker processes of %s; mean inner "
                                   "trips %.1f" % (sample_docs, per_worker, cores, what, trips)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))
    return 0


def cpu_baseline_single_core(cfg, args, row_ptr, ids, cts, eta, alpha, ndocs):
    """The reference's path on ONE host core (the reference is single-threaded), on the first ndocs documents."""
    worker, kind, what = reference_runner(cfg, args)
    _W.update(row_ptr=row_ptr, ids=ids, cts=cts, eta=eta, alpha=alpha, vocab=None)
    if kind == "reference":
        _W["vocab"] = ["w%d" % i for i in range(eta.shape[1])]
    n, dt, trips = worker((0, min(ndocs, len(row_ptr) - 1)))
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "first %d docs of the same corpus, eta0, 1 process / 1 thread, %s; mean inner trips %.1f; %.1f s" % (
                n, what, trips / max(1, n), dt)}


# ------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------
def run_product(args, cfg):
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

    K, V = cfg["K"], cfg["V"]
    row_ptr, ids, cts = rank_corpus(cfg, args.docs, rank, world, args.scaling)
    D = len(row_ptr) - 1
    eta0 = synthetic.initial_eta(K, V, 0)
    alpha = numpy.full(K, 1.0 / K)
    alpha_beta = 1.0 / V
    ctx.set_corpus(0, row_ptr, ids, cts)
    nnz = int(len(ids))

    def barrier():
        if dist is not None:
            dist.barrier()

    def allred(x, op):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t[0])

    def allmax(x):
        return allred(x, dist.ReduceOp.MAX) if dist is not None else x

    def allsum(x):
        return allred(x, dist.ReduceOp.SUM) if dist is not None else x

    nnz0 = nnz
    if dist is not None:                      # the config string quotes rank 0's nnz on every rank
        obj = [nnz if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        nnz0 = obj[0]
    docs_total = allsum(float(D))

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

    em_stats = {}

    em_state = {}

    def warm_model(want_eta, n_docs_total, n_em=4, restart=True):
        """EM iterations 1..4 of variational_bayes.py:239-261 with everything resident: device E-step, device
        M-step, alpha statistics from the device (summed over ranks by the library) and the reference's Newton
        update of alpha on the host (K numbers).  Leaves the model of EM iteration 5 on the device.
        restart=False: n_em further iterations on the model and alpha the previous call left."""
        from pylda_b200.variational_bayes import VariationalBayes
        if restart:
            shell = VariationalBayes()
            shell._number_of_topics = K
            shell._number_of_documents = int(n_docs_total)
            shell._alpha_alpha = alpha.copy()
            ctx.set_model(eta0, alpha)
            em_state["shell"] = shell
        shell = em_state["shell"]
        eta_host = None
        em_wall = []
        for em in range(n_em):
            t_em = time.perf_counter()
            ctx.estep_resident(0, 50, 1e-6, want_alpha_ss=True)
            alpha_ss = ctx.get_results(0, gamma=False, phi=False, alpha_ss=True)["alpha_ss"]
            _, eta_host = ctx.mstep_resident(alpha_beta, want_eta=(want_eta and em == n_em - 1))
            shell.optimize_hyperparameters(alpha_ss)
            ctx.set_alpha(shell._alpha_alpha)
            em_wall.append(time.perf_counter() - t_em)
        em_stats.setdefault("ms", 1e3 * min(em_wall[:3]))      # headline corpus only; the 4th may include the eta copy-back
        return eta_host, shell._alpha_alpha.copy()

    results = {}
    sampler = ClockSampler(local_rank)
    clocks = None
    states = ("cold", "warm") if not args.no_warm else (args.state,)
    for state in states:
        if state == "warm":
            warm_model(False, docs_total)
        else:
            ctx.set_model(eta0, alpha)
        for _ in range(args.warmup):
            ctx.estep_resident(0, 50, 1e-6)
        if state == args.state:
            sampler.start()
        wall, dev, ker, st = timed_resident(args.steps)
        if state == args.state:
            clocks = sampler.stop()
        res = ctx.get_results(0, gamma=False, phi=False)
        results[state] = dict(dev_ms=allmax(sum(dev) / len(dev)), wall_ms=allmax(1e3 * wall / args.steps),
                              ker_ms=allmax(sum(ker) / len(ker)), stats=st, doc_ll=res["doc_ll"],
                              allreduce_ms=allmax(st["allreduce_ms"]),
                              mean_trips=st["inner_iters"] / docs_total, at_cap=st["docs_at_cap"])
        if state == args.state:
            # ---- self-check of the timed state (every N): the two invariants of the E-step ----
            full = ctx.get_results(0, gamma=True, phi=True)
            tokens_total = allsum(float(cts.sum()))
            Nd = numpy.add.reduceat(cts.astype(numpy.float64), row_ptr[:-1]) if D else numpy.zeros(0)
            Nd[numpy.diff(row_ptr) == 0] = 0.0
            want = alpha.sum() + Nd
            check = {
                "what": "sum(phi_ss) vs tokens of ALL ranks (the statistics are all-reduced); max_d |sum_k gamma_dk - "
                        "(sum alpha + N_d)| / (sum alpha + N_d) over all ranks' documents",
                "phi_sum_rel_residual": abs(float(full["phi_ss"].sum()) - tokens_total) / tokens_total,
                "gamma_rowsum_max_rel_residual": allmax(float(numpy.max(numpy.abs(full["gamma"].sum(axis=1) - want) / want))
                                                        if D else 0.0),
                "revived_docs": int(st["revived_docs"]),
            }
            check["ok"] = bool(check["phi_sum_rel_residual"] <= 1e-9 and check["gamma_rowsum_max_rel_residual"] <= 1e-9
                               and check["revived_docs"] == 0)
            del full

    # ---- memory-path probe: the same E-step limited to ONE trip per document (local_parameter_iteration=1):
    # gather + one fixed-point trip + scatter, i.e. the regime where the HBM roof binds ----
    ctx.set_model(eta0, alpha)
    for _ in range(2):
        ctx.estep_resident(0, 1, 1e-6)
    probe_ms = allmax(sorted(ctx.estep_resident(0, 1, 1e-6)["kernel_ms"] for _ in range(3))[1])

    # ---- e2e: the reference-facing call with host buffers (pinned), at the headline state ----
    e2e = None
    alpha_e2e, eta_host = alpha, eta0
    if not args.no_e2e:
        if args.state == "warm":
            eta_host, alpha_e2e = warm_model(True, docs_total)
        eta_pin = numpy.ascontiguousarray(eta_host)
        gamma_pin = numpy.zeros((D, K), dtype=numpy.float64)
        phi_pin = numpy.zeros((K, V), dtype=numpy.float64)
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
        e2e = {"value": docs_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": eta_pin.nbytes + alpha.nbytes,
               "d2h_bytes_per_step": gamma_pin.nbytes + phi_pin.nbytes + 64, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "elbo_doc_ll": out["doc_ll"],
               "host_buffers": "page-locked (cudaHostRegister): eta H2D and phi_ss D2H by cudaMemcpyAsync; the gamma D2H starts "
                               "when the short documents are final and overlaps the long-document kernels"}
        for a in (eta_pin, gamma_pin, phi_pin):
            ctx.unpin(a)
        del gamma_pin, phi_pin

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
    fp64_ach = 4.0 * K * st["row_trips"] / max(1, world) / (head["ker_ms"] * 1e-3) / 1e12      # row_trips: summed over ranks

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_single_core(cfg, args, row_ptr, ids, cts, eta_host, alpha_e2e, args.cpu_docs)

    # ---- warm_lda: EM iteration 5 on a corpus drawn from an actual LDA model (the fixed point then stops after
    # ~10 trips: the regime where the HBM roof is the one that binds) ----
    warm_lda = None
    if not args.no_warm_lda:
        lr, li, lc = load_corpus(args.lda_docs, V, 4321 + rank, kind="lda")
        ctx.set_corpus(0, lr, li, lc)
        lda_total = allsum(float(len(lr) - 1))
        warm_model(False, lda_total)
        for _ in range(2):
            ctx.estep_resident(0, 50, 1e-6)
        _, ldev, lker, lst = timed_resident(3)
        lms, lk = allmax(sum(ldev) / 3), allmax(sum(lker) / 3)
        warm_lda = {"what": "EM iteration 5 on %d documents per GPU drawn from an LDA model (50 topics, "
                            "pylda_b200.synthetic.lda_corpus), same V, K and Zipf lengths" % (len(lr) - 1),
                    "value": lda_total / (lms * 1e-3), "unit": UNIT, "ms_per_step": lms, "kernel_ms": lk,
                    "mean_inner_trips": lst["inner_iters"] / lda_total,
                    "roofline_frac": lst["algo_total_bytes"] / (lk * 1e-3) / 1e9 / peak,
                    "roofline_frac_read": lst["algo_read_bytes"] / (lk * 1e-3) / 1e9 / peak}
        # ... and later in the same training run (EM iteration 20), where fewer trips are left per document
        warm_model(False, lda_total, n_em=15, restart=False)
        for _ in range(2):
            ctx.estep_resident(0, 50, 1e-6)
        _, ldev, lker, lst = timed_resident(3)
        lms, lk = allmax(sum(ldev) / 3), allmax(sum(lker) / 3)
        warm_lda["em_iteration_20"] = {"value": lda_total / (lms * 1e-3), "unit": UNIT, "ms_per_step": lms, "kernel_ms": lk,
                                       "mean_inner_trips": lst["inner_iters"] / lda_total,
                                       "roofline_frac": lst["algo_total_bytes"] / (lk * 1e-3) / 1e9 / peak}

    if rank == 0:
        traffic, traffic_note = measured_traffic(args.config, args.docs)
        line = {
            "metric": cfg["metric"], "value": docs_total / (head["dev_ms"] * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["dev_ms"],
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": shared_config(args, cfg, nnz0),
            "run": {"mean_inner_trips": head["mean_trips"], "docs_at_cap": head["at_cap"], "docs_total": docs_total,
                    "l2": "inputs larger than L2 (CSR + gamma + tables = %.1f GB per GPU)" % (
                        (12.0 * nnz + 8.0 * D * (K + 2) + 4 * 8.0 * V * K) / 1e9),
                    "parallelism": "dp%d, one process per GPU, one NCCL all-reduce of K x V f64 per step" % world,
                    "timing": "CUDA events on the library's stream, per step, max over ranks",
                    "docs_narrow_stages": [int(st["docs_narrow_wide"]), int(st["docs_narrow"])],
                    "allreduce_ms": head["allreduce_ms"]},
            "wall_ms_per_step": head["wall_ms"],
            "elbo_doc_ll": head["doc_ll"],
            "check": check,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                         "kernel": "per-document E-step kernels of one E-step (register-tile, shared-memory, streaming and "
                                   "narrow-stage kernels: one launch per length class / stage)",
                         "kernel_ms": head["ker_ms"], "achieved_read": achieved_read,
                         "frac_read": achieved_read / peak,
                         "note": "binding roof at %.1f trips/doc is the fp64 pipe / on-chip latency, not HBM (DESIGN.md 4); "
                                 "see roofline_fp64 and warm_lda" % head["mean_trips"]},
            "roofline_fp64": {"bound": "fp64 pipe", "achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": fp64_ach / fp64_peak,
                              "flops": "4*K per (term row, trip) of the two mat-vecs as the reference performs them (the "
                                       "kernels skip eliminated topics: an effective rate)",
                              "peak_source": "measured DFMA rate, scripts/ubench/fp64_lat.cu (profiles/r1_fp64_ubench.txt)"},
            "em_iteration": {"what": "one whole resident EM iteration as VariationalBayes.learning() runs it: E-step + alpha "
                                     "statistics + device M-step + host Newton update of alpha (wall clock, rank 0)",
                             "ms": em_stats.get("ms")},
            "probe_1trip": {"what": "same corpus with local_parameter_iteration=1 (gather + one trip + scatter): the "
                                    "regime where the HBM roof binds", "kernel_ms": probe_ms,
                            "achieved": algo / (probe_ms * 1e-3) / 1e9, "unit": "GB/s",
                            "frac": algo / (probe_ms * 1e-3) / 1e9 / peak},
            "gpu_launches": int(st["n_launches"]) * args.steps,
            "estep_kernel_launches_per_step": int(st["n_estep_launches"]),
            "docs_resident": st["docs_resident"], "docs_streamed": st["docs_streamed"],
            "clocks": clocks,
            "device": ctx.device_name(),
            "build_digest": build_digest(),
        }
        other = "warm" if args.state == "cold" else "cold"
        if other in results:
            o = results[other]
            line[other] = {"value": docs_total / (o["dev_ms"] * 1e-3), "ms_per_step": o["dev_ms"], "kernel_ms": o["ker_ms"],
                           "mean_inner_trips": o["mean_trips"], "elbo_doc_ll": o["doc_ll"],
                           "roofline_frac": o["stats"]["algo_total_bytes"] / (o["ker_ms"] * 1e-3) / 1e9 / peak,
                           "roofline_frac_read": o["stats"]["algo_read_bytes"] / (o["ker_ms"] * 1e-3) / 1e9 / peak}
        if warm_lda is not None:
            line["warm_lda"] = warm_lda
        if e2e is not None:
            line["e2e"] = e2e
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
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--docs", type=int, default=None, help="documents per GPU (weak) or in total (strong)")
    ap.add_argument("--state", default="cold", choices=["cold", "warm"])
    ap.add_argument("--cpu-docs", type=int, default=2000)
    ap.add_argument("--ref-docs-per-core", type=int, default=150)
    ap.add_argument("--ref-port", action="store_true", help="reference arm: time the numpy restatement even when the reference is present")
    ap.add_argument("--lda-docs", type=int, default=250000, help="documents per GPU of the warm_lda measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-warm", action="store_true", help="skip the other state (warm when --state cold)")
    ap.add_argument("--no-warm-lda", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.docs is None:
        args.docs = cfg["D"]
    if args.config not in ("c2", "c3"):          # the K = 500 configs: 4 GB tables; keep the run bounded
        args.no_warm = True
        args.no_warm_lda = True
    if args.warmup < 3 and args.impl == "b200":
        log("note: W >= 3 warm-up steps are required for a valid number; got %d" % args.warmup)
    if args.impl == "reference":
        return run_reference(args, cfg)
    return run_product(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
