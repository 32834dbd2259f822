"""cluster_log.csv / job_log.csv wire formats (SIM:414-421, 815-823, 944-948) against files written by the
UNMODIFIED reference (tests/golden/csv/, made by tests/golden/make_golden_csv.py with the Philox stream injected).

CPU: the product's CSV writer fed with the oracle's unrounded records must reproduce the reference files byte for
byte.  GPU (-m gpu): the drop-in simulator's own files must have identical keys/integers and numerically equal
floats (the device's libm may differ from glibc in the last ulp, which a rounded field can expose)."""
import csv
import filecmp
import logging
import os

import pytest

from conftest import GOLDEN_DIR, has_cuda
from distributed_cluster_gpus_b200 import scenarios as SC
from distributed_cluster_gpus_b200.simcore.simulator_paper_multi import write_csv_logs

CASES = [("ragged_3dc_12_5_40", 123), ("csv_joint_nf_4x64_20s", 7), ("csv_carbon_cost_2x16", 11)]
STOCK_CASES = [("ragged_3dc_12_5_40", 123), ("csv_carbon_cost_2x16", 42)]   # the reference exactly as shipped (its own MT19937)


def golden(name, seed, which):
    return os.path.join(GOLDEN_DIR, "csv", f"{name}_seed{seed}_{which}_log.csv")


@pytest.mark.parametrize("name,seed", CASES)
def test_writer_reproduces_reference_files_byte_for_byte(oracle, tmp_path, name, seed):
    sc = SC.CSV_SCENARIOS[name]
    kw = SC.build_inputs(sc)
    sp = SC.to_spec(sc)
    sim = oracle.OracleSim(sp.to_bytes(), seed, joblog_cap=100000, clog_cap=10000)
    sim.advance(0)
    cl, jl = str(tmp_path / "cluster_log.csv"), str(tmp_path / "job_log.csv")
    write_csv_logs(sim.job_log(), sim.cluster_log(), kw["dcs"], list(kw["ingresses"]), kw["coeffs_map"], sp.net_lat_s, cl, jl)
    assert filecmp.cmp(jl, golden(name, seed, "job"), shallow=False)
    assert filecmp.cmp(cl, golden(name, seed, "cluster"), shallow=False)


@pytest.mark.parametrize("name,seed", STOCK_CASES)
def test_device_core_in_mt_mode_reproduces_stock_reference_files(hostemu, tmp_path, name, seed):
    """No Philox anywhere: the device core (host build) drawing from CPython's Mersenne Twister + the product's CSV
    writer == the files `python run_sim_paper.py` of the untouched reference writes at the same seed."""
    from distributed_cluster_gpus_b200.engine import CLUSTER_DTYPE, JOB_DTYPE
    sc = SC.CSV_SCENARIOS[name]
    kw = SC.build_inputs(sc)
    sp = SC.to_spec(sc)
    got = hostemu.run_batch(sp.to_bytes(), 1, seed, rec_replica=0, job_dtype=JOB_DTYPE, jobs_cap=100000,
                            cluster_dtype=CLUSTER_DTYPE, cluster_cap=10000, rng_kind=1)
    cl, jl = str(tmp_path / "cluster_log.csv"), str(tmp_path / "job_log.csv")
    write_csv_logs(got["jobs"], got["cluster"], kw["dcs"], list(kw["ingresses"]), kw["coeffs_map"], sp.net_lat_s, cl, jl)
    assert filecmp.cmp(jl, golden(name, f"{seed}_mt", "job"), shallow=False)
    assert filecmp.cmp(cl, golden(name, f"{seed}_mt", "cluster"), shallow=False)


def _rows(path):
    with open(path) as f:
        return list(csv.reader(f))


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("name,seed", CASES + [(n, f"{s}_mt") for n, s in STOCK_CASES])
def test_drop_in_csvs_match_reference_files(tmp_path, name, seed):
    from distributed_cluster_gpus_b200.configs import paper_config as pc
    from distributed_cluster_gpus_b200.simcore.simulator_paper_multi import MultiIngressPaperSimulator
    sc = SC.CSV_SCENARIOS[name]
    kw = SC.build_inputs(sc)
    stock = isinstance(seed, str)          # "<seed>_mt": files of the reference as shipped -> rng="mt19937"
    MultiIngressPaperSimulator(router_policy=pc.build_router_policy(), logger=logging.getLogger("t"),
                               sim_duration=sc["duration"], log_interval=sc["log_interval"], log_path=str(tmp_path),
                               rng_seed=int(seed.split("_")[0]) if stock else seed, algo=sc["algo"], power_cap=sc["power_cap"],
                               show_progress=False, num_fixed_gpus=sc["num_fixed_gpus"], fixed_freq=sc["fixed_freq"],
                               replicas=4, rng="mt19937" if stock else "philox", **kw).run()
    for which, textual in (("job", {1, 2, 4}), ("cluster", {1})):
        got, want = _rows(tmp_path / f"{which}_log.csv"), _rows(golden(name, seed, which))
        assert got[0] == want[0] and len(got) == len(want)
        exact = 0
        for g, w in zip(got[1:], want[1:]):
            assert len(g) == len(w)
            for k, (a, b) in enumerate(zip(g, w)):
                if k in textual or a == b:
                    assert a == b
                    exact += 1
                else:   # a rounded field whose last printed digit flipped on a last-ulp difference
                    decimals = len(b.split(".")[1]) if "." in b else 0
                    assert abs(float(a) - float(b)) <= 1.01 * 10 ** (-decimals), (which, k, a, b)
        assert exact >= 0.999 * sum(len(r) for r in want[1:])


def _cli_cases():
    import json
    with open(os.path.join(GOLDEN_DIR, "cli", "cases.json")) as f:
        return sorted(json.load(f).items())


@pytest.mark.parametrize("name,flags", _cli_cases())
def test_cli_configuration_reproduces_the_reference_command_line(hostemu, tmp_path, name, flags):
    """`python run_sim_paper.py <flags>` of the untouched reference wrote tests/golden/cli/*.  The product's CLI builds
    its simulator from the same flags; its spec through the device core (host build, rng = MT19937) and its CSV
    writer must give the same bytes — scenario defaults, builders, flattening and writer all in one comparison."""
    from distributed_cluster_gpus_b200 import run_sim_paper as cli
    from distributed_cluster_gpus_b200.engine import CLUSTER_DTYPE, JOB_DTYPE
    args = cli.parse_args(flags + ["--log-path", str(tmp_path / "logs" / "x"), "--rng", "mt19937"])
    sim = cli.build_simulator(args)
    got = hostemu.run_batch(sim._spec.to_bytes(), 1, args.seed, rec_replica=0, job_dtype=JOB_DTYPE, jobs_cap=200000,
                            cluster_dtype=CLUSTER_DTYPE, cluster_cap=20000, rng_kind=1)
    assert int(got["summary"][0, 0]) == 0
    sim._write_csvs(got["jobs"], got["cluster"])
    for f in ("cluster_log.csv", "job_log.csv"):
        assert filecmp.cmp(str(tmp_path / "logs" / "x" / f), os.path.join(GOLDEN_DIR, "cli", f"{name}_{f}"), shallow=False), f


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("name,flags", _cli_cases())
def test_cli_on_device_matches_the_reference_command_line(tmp_path, name, flags):
    """The product's command line, end to end on the GPU, against the files the reference's command line wrote."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    logs = tmp_path / "logs" / "x"
    subprocess.run([sys.executable, "-m", "distributed_cluster_gpus_b200.run_sim_paper", "--log-path", str(logs), "--rng", "mt19937",
                    "--replicas", "3"] + flags, cwd=root, check=True, capture_output=True, timeout=600)
    for which, textual in (("job", {1, 2, 4}), ("cluster", {1})):
        got, want = _rows(logs / f"{which}_log.csv"), _rows(os.path.join(GOLDEN_DIR, "cli", f"{name}_{which}_log.csv"))
        assert got[0] == want[0] and len(got) == len(want)
        for g, w in zip(got[1:], want[1:]):
            assert len(g) == len(w)
            for k, (a, b) in enumerate(zip(g, w)):
                if k in textual or a == b:
                    assert a == b
                else:
                    decimals = len(b.split(".")[1]) if "." in b else 0
                    assert abs(float(a) - float(b)) <= 1.01 * 10 ** (-decimals), (which, k, a, b)
