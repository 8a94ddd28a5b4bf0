"""RHLV feature table and SVM Genant grading: the tail of the pipeline.

* ``rhlv_table``        - reference ``process_datasets_to_excel`` (evaluation/RHLV_quantification.py:150-195; coronal twin with
                          ``axis=1``): per vertebra volume, window ``[c - len, c + len)`` around the label's centre slice,
                          RHLV all / pre / mid / post + relative label height.  The integer column scan runs on the GPU
                          (``mask_ops.calculate_rhlv`` -> ``hv_column_heights``), the float64 ratios on the host; volumes are read with
                          ``healthivert_gan_b200.nifti``.
* ``write_table`` / ``read_table`` - the reference writes ``.xlsx`` through pandas + openpyxl; ``.csv`` is always available here,
                          ``.xlsx`` when openpyxl is importable.
* ``evaluate_svm`` / ``evaluate_svm_25d`` - reference evaluation/SVM_grading.py:9-79 and SVM_grading_2.5d.py:9-82: StandardScaler on
                          train+test rows, linear ``SVC(class_weight='balanced')``, ``StratifiedKFold(5)``, every fold scored on the
                          ``val`` rows; the text report has the reference's layout.  scikit-learn stays on the host (tiny problem).
"""
import csv
import os

import numpy as np

from . import mask_ops, nifti

COLUMNS = ["Vertebra", "Label", "Dataset", "All RHLV", "Pre RHLV", "Mid RHLV", "Post RHLV", "Relative Height Label"]


def rhlv_row(vertebra, label, dataset, label_volume, fake_volume, length_divisor=5, height_threshold=0.64, axis=2):
    """One row of the table from two label maps (arrays); None when the vertebra id is absent from the label map
    (`continue` at RHLV_quantification.py:171-172)."""
    label_index = int(vertebra.split("_")[-1])
    seg_label = (np.asarray(label_volume) == label_index).astype(np.float64)
    seg_fake = (np.asarray(fake_volume) == label_index).astype(np.float64)
    loc = np.where(seg_label)[axis]
    if loc.size == 0:
        return None
    min_z, max_z = int(np.min(loc)), int(np.max(loc))
    center_z = int(np.mean(loc))
    length = (max_z - min_z) // length_divisor
    a, p, m, q, rel = mask_ops.calculate_rhlv(seg_fake, seg_label, center_z, length, vertebra, height_threshold, axis=axis)
    return {"Vertebra": vertebra, "Label": label, "Dataset": dataset, "All RHLV": a, "Pre RHLV": p, "Mid RHLV": m, "Post RHLV": q,
            "Relative Height Label": rel}


def rhlv_table(dataset_info, label_folder, fake_folder, length_divisor=5, height_threshold=0.64, axis=2):
    """dataset_info: {"train" | "test" | "val": {"<patient>_<vert>": grade}} (the reference's vertebra_data.json).  Volumes missing
    on either side are skipped like the reference does."""
    rows = []
    for dataset_type, data in dataset_info.items():
        for vertebra, label in data.items():
            label_path = os.path.join(label_folder, vertebra + ".nii.gz")
            fake_path = os.path.join(fake_folder, vertebra + ".nii.gz")
            if not os.path.exists(label_path) or not os.path.exists(fake_path):
                continue
            row = rhlv_row(vertebra, label, dataset_type, nifti.load(label_path).get_fdata(), nifti.load(fake_path).get_fdata(),
                           length_divisor, height_threshold, axis)
            if row is not None:
                rows.append(row)
    return rows


def write_table(rows, path):
    if str(path).endswith(".xlsx"):
        import pandas as pd   # needs openpyxl, like the reference
        pd.DataFrame(rows, columns=COLUMNS).to_excel(path, index=False)
        return
    with open(path, "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=COLUMNS)
        w.writeheader()
        for r in rows:
            w.writerow({k: (repr(float(v)) if isinstance(v, (float, np.floating)) else v) for k, v in r.items()})


def read_table(path):
    """-> pandas.DataFrame with the reference's column names (csv, or xlsx when openpyxl is installed)."""
    import pandas as pd
    return pd.read_excel(path) if str(path).endswith((".xlsx", ".xls")) else pd.read_csv(path, float_precision="round_trip")


def _as_frame(table):
    import pandas as pd
    if isinstance(table, pd.DataFrame):
        return table
    if isinstance(table, (str, os.PathLike)):
        return read_table(table)
    return pd.DataFrame(list(table), columns=COLUMNS)


def _cross_validate(X_train_test, y_train_test, X_val, y_val):
    from sklearn.metrics import accuracy_score, confusion_matrix, f1_score, precision_score, recall_score
    from sklearn.model_selection import StratifiedKFold
    from sklearn.preprocessing import StandardScaler
    from sklearn.svm import SVC
    scaler = StandardScaler()
    Xs = scaler.fit_transform(X_train_test)
    Xv = scaler.transform(X_val)
    y = np.asarray(y_train_test)
    clf = SVC(kernel="linear", class_weight="balanced")
    results = []
    for train_index, _ in StratifiedKFold(n_splits=5).split(Xs, y):
        clf.fit(Xs[train_index], y[train_index])
        pred = clf.predict(Xv)
        results.append((confusion_matrix(y_val, pred), f1_score(y_val, pred, average="macro"), precision_score(y_val, pred, average="macro"),
                        recall_score(y_val, pred, average="macro"), accuracy_score(y_val, pred)))
    return results


def evaluate_svm(table, features, output_txt=None):
    """reference evaluation/SVM_grading.py:9-79.  Returns {"folds": [(cm, f1, precision, recall, accuracy)], "mean": {...}, "var": {...}}."""
    data = _as_frame(table)
    tt = data[data["Dataset"].isin(["train", "test"])]
    val = data[data["Dataset"] == "val"]
    results = _cross_validate(tt[features], tt["Label"], val[features], val["Label"])
    names = ["F1 Score", "Precision", "Recall", "Accuracy"]
    cols = [[r[i + 1] for r in results] for i in range(4)]
    mean = {n: np.mean(c) for n, c in zip(names, cols)}
    var = {n: np.var(c) for n, c in zip(names, cols)}
    if output_txt:
        with open(output_txt, "w") as file:
            for i, (cm, f1, precision, recall, accuracy) in enumerate(results):
                file.write(f"Fold {i+1}:\n")
                file.write("Confusion Matrix:\n")
                file.write(f"{cm}\n")
                file.write(f"F1 Score: {f1}, Precision: {precision}, Recall: {recall}, Accuracy: {accuracy}\n")
                file.write("\n")
            file.write("Average Scores:\n")
            for n in names:
                file.write(f"Average {n}: {mean[n]} (Variance: {var[n]})\n")
    return {"folds": results, "mean": mean, "var": var}


def evaluate_svm_25d(table1, table2, features, output_txt=None):
    """reference evaluation/SVM_grading_2.5d.py:9-82: sagittal and coronal tables merged on "Vertebra", features of the second
    table suffixed `_2`."""
    import pandas as pd
    data1, data2 = _as_frame(table1), _as_frame(table2).copy()
    data2.rename(columns={f: f"{f}_2" for f in features}, inplace=True)
    combined = pd.merge(data1, data2, on="Vertebra")
    tt = combined[combined["Dataset_x"].isin(["train", "test"])]
    val = combined[combined["Dataset_x"] == "val"]
    feats = list(features) + [f"{f}_2" for f in features]
    results = _cross_validate(tt[feats], tt["Label_x"], val[feats], val["Label_x"])
    names = ["F1 Score", "Precision", "Recall", "Accuracy"]
    mean = {n: np.mean([r[i + 1] for r in results]) for i, n in enumerate(names)}
    if output_txt:
        with open(output_txt, "w") as file:
            for i, (cm, f1, precision, recall, accuracy) in enumerate(results):
                file.write(f"Fold {i+1}:\n")
                file.write("Confusion Matrix:\n")
                file.write(f"{cm}\n")
                file.write(f"F1 Score: {f1:.3f}, Precision: {precision:.3f}, Recall: {recall:.3f}, Accuracy: {accuracy:.3f}\n")
                file.write("\n")
            file.write("Average Scores:\n")
            for n in names:
                file.write(f"Average {n}: {mean[n]:.3f}\n")
    return {"folds": results, "mean": mean}
