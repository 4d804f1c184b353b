"""`sample_result.txt` -> CSV (reference `uformerWM/result_extract.py:12-39`): same line pattern, same
column names.  PESQ is 'N/A' in lines written by this repo (pypesq is a third-party codec-style
dependency with no source in the reference tree) and is carried through as an empty cell."""
import csv
import re

PATTERN = (r"Result on (.*), attack: (.*): Total clips: (.*), MSE loss (.*), WM loss: (.*), "
           r"WM loss after attack: (.*), SNR score: (.*), PESQ score: (.*)")
FIELDNAMES = ["Set", "Attack", "Total Clips", "MSE Loss", "WM Loss", "WM Loss After Attack", "SNR Score", "PESQ Score"]


def _num(s):
    try:
        return float(s)
    except ValueError:
        return ""


def parse_results(text):
    """List of row dicts for every result line in `text` (`result_extract.py:14-28`)."""
    rows = []
    for r in re.findall(PATTERN, text):
        rows.append({"Set": r[0], "Attack": r[1], "Total Clips": int(float(r[2])), "MSE Loss": float(r[3]),
                     "WM Loss": float(r[4]), "WM Loss After Attack": float(r[5]), "SNR Score": float(r[6]),
                     "PESQ Score": _num(r[7].strip())})
    return rows


def process_data_to_csv(data, csv_path):
    """Write the CSV table (`result_extract.py:30-36`); returns the rows."""
    rows = parse_results(data)
    with open(csv_path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=FIELDNAMES)
        w.writeheader()
        for row in rows:
            w.writerow(row)
    return rows
