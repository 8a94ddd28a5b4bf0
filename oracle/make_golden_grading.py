"""Generate tests/golden/grading.json by running the UNMODIFIED reference grading scripts (build container only).

TEST INFRASTRUCTURE.  Run from the repo root:  python -m oracle.make_golden_grading
evaluation/SVM_grading.py and SVM_grading_2.5d.py read .xlsx files through pandas + openpyxl (absent here): pd.read_excel is
replaced by a function that hands them the synthetic feature tables below; everything else is the reference's own code.  The
fixture holds the tables and the text reports the reference wrote.
"""
import importlib.util
import json
import os
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FEATURES = ["Pre RHLV", "Mid RHLV", "Post RHLV"]


def synthetic_table(seed):
    """Three Genant-like grades; RHLV features grow with the grade, with overlap.  Rows ordered train, test, val like the reference's
    process_datasets_to_excel output."""
    rng = np.random.default_rng(seed)
    rows = []
    for dataset, count in (("train", 45), ("test", 15), ("val", 24)):
        for i in range(count):
            grade = int(rng.integers(0, 3))
            base = 0.08 + 0.12 * grade
            f = np.clip(base + rng.normal(0, 0.06, 3) + np.array([0.0, 0.04 * grade, 0.0]), -0.2, 0.9)
            rows.append({"Vertebra": f"{dataset}{i:03d}_{17 + i % 7}", "Label": grade, "Dataset": dataset, "All RHLV": float(f.mean()),
                         "Pre RHLV": float(f[0]), "Mid RHLV": float(f[1]), "Post RHLV": float(f[2]),
                         "Relative Height Label": float(np.clip(1.0 - 0.15 * grade + rng.normal(0, 0.03), 0, 1))})
    return rows


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    t1, t2 = synthetic_table(11), synthetic_table(12)
    for a, b in zip(t1, t2):                      # the coronal table describes the same vertebrae
        b["Vertebra"], b["Label"], b["Dataset"] = a["Vertebra"], a["Label"], a["Dataset"]
    frames = {"one.xlsx": pd.DataFrame(t1), "two.xlsx": pd.DataFrame(t2)}
    pd.read_excel = lambda p, *a, **k: frames[os.path.basename(p)].copy()
    g1 = _load("/root/reference/evaluation/SVM_grading.py", "ref_svm_grading")
    g2 = _load("/root/reference/evaluation/SVM_grading_2.5d.py", "ref_svm_grading_25d")
    with tempfile.TemporaryDirectory() as d:
        g1.evaluate_svm("one.xlsx", FEATURES, os.path.join(d, "a.txt"))
        g2.evaluate_svm("one.xlsx", "two.xlsx", FEATURES, os.path.join(d, "b.txt"))
        rep1, rep2 = open(os.path.join(d, "a.txt")).read(), open(os.path.join(d, "b.txt")).read()
    with open(os.path.join(ROOT, "tests", "golden", "grading.json"), "w") as fh:
        json.dump({"features": FEATURES, "table1": t1, "table2": t2, "report_single": rep1, "report_25d": rep2}, fh, indent=0)
    print(rep1[-400:])
    print(rep2[-300:])


if __name__ == "__main__":
    main()
