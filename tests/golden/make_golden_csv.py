"""Golden cluster_log.csv / job_log.csv written by the UNMODIFIED reference, for the CSV wire-format tests: with the
Philox stream injected (CASES) and exactly as shipped, on its own Mersenne Twister (STOCK_CASES, files *_mt_*).
Build-container only.   python tests/golden/make_golden_csv.py"""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from distributed_cluster_gpus_b200 import scenarios as S  # noqa: E402
from ref_harness import run_reference  # noqa: E402

CASES = [("ragged_3dc_12_5_40", 123), ("csv_joint_nf_4x64_20s", 7), ("csv_carbon_cost_2x16", 11)]
STOCK_CASES = [("ragged_3dc_12_5_40", 123), ("csv_carbon_cost_2x16", 42)]

if __name__ == "__main__":
    out = os.path.join(HERE, "csv")
    os.makedirs(out, exist_ok=True)
    for name, seed in CASES:
        sc = S.CSV_SCENARIOS[name]
        with tempfile.TemporaryDirectory() as tmp:
            run_reference(sc, seed, rng="philox", log_dir=tmp)
            for f in ("cluster_log.csv", "job_log.csv"):
                shutil.copy(os.path.join(tmp, f), os.path.join(out, f"{name}_seed{seed}_{f}"))
        print(name, seed, "ok")
    for name, seed in STOCK_CASES:
        sc = S.CSV_SCENARIOS[name]
        with tempfile.TemporaryDirectory() as tmp:
            run_reference(sc, seed, rng="mt", log_dir=tmp)
            for f in ("cluster_log.csv", "job_log.csv"):
                shutil.copy(os.path.join(tmp, f), os.path.join(out, f"{name}_seed{seed}_mt_{f}"))
        print(name, seed, "stock (mt) ok")
