"""Write thepayne_b200/data/highav_coeffs.txt from a reference checkout (dev tool, run once).

The per-band coefficients of the reference's Av >= 5 extension (Payne/predict/highred.py:29-169)
are data, not code: this script pulls the rows `filter a1 b1 a2 b2 c2` out of the reference file
and writes them as a whitespace table that thepayne_b200.predict.highred.highAv reads at run time.

    python tools/export_highav.py [/root/reference/Payne/predict/highred.py]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thepayne_b200.predict.highred import highAv, TABLE_PATH  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/Payne/predict/highred.py'
print(highAv.export_reference_table(src, TABLE_PATH))
print(sum(1 for _ in open(TABLE_PATH)) - 1, 'bands')
