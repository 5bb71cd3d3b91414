"""bench.py's JSON contract: the reference arm run here on a tiny sample (CPU only), and the last
committed GPU-arm line under profiles/ (written by `python bench.py` on a B200)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "cpu_baseline"]


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "ecoli100x", "--cpu-sample-reads", "3000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    for k in BASE:
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "input bases/sec to finished seqset" and line["unit"] == "bases/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 0
    assert line["config"]["workload"] == "ecoli100x"
    # the same `config` dict as the GPU arm prints for this job (r2l_bench_ecoli100x.json: 3 292 614 reads of 150 bases)
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.arm_config("ecoli100x", 3292614, 150, 1) and "host threads" in line["host_parallelism"]
    assert line["e2e"] == {"value": line["value"], "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    # the reference's own classes where oracle/_ref was built (development container; travels to the GPU box), else the port
    from oracle import ref as R
    assert cb["kind"] == ("reference" if R.available() else "port") and cb["value"] == line["value"] and "sample" in cb
    # all the host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
    assert cb["cores"] == len(os.sched_getaffinity(0))


def test_reference_arm_ignores_omp_num_threads():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "small", "--cpu-sample-reads", "2000"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["config"]["workload"] == "small"


@pytest.mark.parametrize("name,workload,n_gpus", [("r1g_bench_ecoli100x.json", "ecoli100x", 1), ("r2l_bench_chr20.json", "chr20_30x", 1),
                                                  ("r2n_bench_chr20_8gpu.json", "chr20_30x", 8)])
def test_committed_gpu_line_has_every_contract_key(name, workload, n_gpus):
    path = os.path.join(ROOT, "profiles", name)
    line = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    for k in BASE + ["gpu_launches", "roofline", "clocks"]:
        assert k in line, k
    assert line["n_gpus"] == n_gpus and line["warmup"] >= 3 and line["gpu_launches"] > 0 and line["dtype"] == "u64"
    assert line["config"]["workload"] == workload and "l2" in line["config"]
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < line["value"]
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    if n_gpus == 1:   # the CPU arm runs on rank 0 at N=1 only
        cb = line["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0
    c = line["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if name.startswith("r2"):   # round 2: the whole workload is checked inside the bench run
        p = line["parity"]
        assert p["checked"] and p["members_equal"] and p["mismatches"] == [] and p["entries"] > 100_000_000
        assert ("oracle" in p["against"]) == (n_gpus == 1)
        if n_gpus == 1:
            assert line["cpu_baseline"]["same_config"] is True


def test_sample_check_against_the_reference_with_a_stand_in_device():
    """bench.py's second parity anchor (a GPU build of the cpu_baseline sample against oracle/_ref) -- its comparison
    logic, run here with the oracle standing in for the device (the real thing needs a B200)."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle as O
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built")

    class FakeBgx:
        def __init__(self, device=0):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            pass

        def add_reads(self, reads):
            buf, offs = reads
            self.rb = (buf.tobytes(), offs)

        def run(self):
            self.solid = O.solid_set(O.count_kmers(self.rb, 30), 5)
            self.cr = O.correct_reads(self.rb, self.solid, 30)
            self.ss = O.seqset_staged((self.cr["seq"], self.cr["offs"]), self.cr["next_fwd"], self.cr["next_rev"])

        def export_kmers(self, min_count):
            return self.solid

        def export_corrected(self):
            return self.cr

        def export_seqset(self):
            return self.ss

    class FakeB:
        Bgx = FakeBgx

    sub = bench.make_workload("small", 0, None, genome_prefix_reads=4000)
    # as the bench does it: the reference runs in a child process that regenerates the same sample
    dt, n_ent, stage, res = bench.cpu_ref_run_isolated("small", sub.shape[0], 2, bench.WORKLOADS["small"]["coverage"])
    same = bench.cpu_ref_run(sub, 2, bench.WORKLOADS["small"]["coverage"], keep=True)[3]
    assert same["corrected"]["seq"] == res["corrected"]["seq"] and np.array_equal(same["seqset"]["prev"], res["seqset"]["prev"])
    par = bench.verify_sample_against_reference(FakeB, 0, sub, res)
    assert par["checked"] and par["members_equal"] and par["mismatches"] == [] and par["entries"] == n_ent > 0
    # a wrong table is reported, not raised
    res["seqset"]["shared"] = res["seqset"]["shared"].copy()
    res["seqset"]["shared"][5] += 1
    par = bench.verify_sample_against_reference(FakeB, 0, sub, res)
    assert par["checked"] and not par["members_equal"] and par["mismatches"] == ["seqset/shared"]


@pytest.mark.parametrize("workload,gpu_line,ranks", [("ecoli100x", "r2l_bench_ecoli100x.json", 1), ("chr20_30x", "r2l_bench_chr20.json", 1),
                                                     ("ecoli100x", "r2k_bench_ecoli100x_2gpu.json", 2),
                                                     ("chr20_30x", "r2k_bench_chr20_2gpu.json", 2)])
def test_gpu_line_digests_equal_the_reference_over_the_whole_workload(workload, gpu_line, ranks):
    """Two committed records that never met on one machine: the bench line measured on a B200 (its `parity.sha256_16`
    are digests of what the CUDA path produced for the whole workload) and tools/ref_full_workload.py's digests of what
    the reference's OWN classes (oracle/_ref) produce for the same workload on a CPU.  Equal digests = equal bytes."""
    sys.path.insert(0, ROOT)
    import bench
    # ranks = 2: one sharded build on two GPUs over two genomes' reads; its line digests the assembled tables
    ref_file = os.path.join(ROOT, "profiles", f"r2u_oracle_vs_reference_{workload}" + (f"_x{ranks}" if ranks > 1 else "") + ".json")
    if not os.path.exists(ref_file) or os.path.getsize(ref_file) == 0:
        pytest.skip("no committed reference digests for " + workload)
    rj = json.load(open(ref_file))
    assert rj["equal"] in (True, None) and rj["mismatches"] == []  # the oracle port agreed with the reference there (None: not run)
    line = json.loads([l for l in open(os.path.join(ROOT, "profiles", gpu_line)) if l.startswith("{")][-1])
    assert line["config"]["workload"] == workload and line["parity"]["members_equal"] is True
    got = bench.reference_full_workload_digests(workload, line["parity"], ranks)
    assert got["digests_equal"] is True and got["differing"] == [] and got["digests_compared"] == (18 if ranks == 1 else 15)
    assert got["entries"] == line["parity"]["entries"] == rj["entries"]
    # and a changed member is noticed
    bad = json.loads(json.dumps(line["parity"]))
    key = "seqset/prev_G/bits" if ranks == 1 else "prev_G/bits"
    bad["sha256_16"][key] = "0" * 16
    assert bench.reference_full_workload_digests(workload, bad, ranks)["differing"] == [key]
