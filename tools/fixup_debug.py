import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
from oracle import pcd_oracle as O
F = pcd.functional
FORMS = {"row_col_mulsum": (F.FORM_ROW_COL, F.NORM_MULSUM, O.FORM_ROW_COL, O.NORM_MULSUM),
         "col_row_mulsum": (F.FORM_COL_ROW, F.NORM_MULSUM, O.FORM_COL_ROW, O.NORM_MULSUM),
         "sum_first_fma": (F.FORM_SUM_FIRST, F.NORM_FMA, O.FORM_SUM_FIRST, O.NORM_FMA),
         "row_col_fma": (F.FORM_ROW_COL, F.NORM_FMA, O.FORM_ROW_COL, O.NORM_FMA),
         "sum_first_mulsum": (F.FORM_SUM_FIRST, F.NORM_MULSUM, O.FORM_SUM_FIRST, O.NORM_MULSUM)}
B, N, M = 3, 700, 1000
rs = np.random.RandomState(B * 7919 + N * 31 + M)
cols = (rs.rand(B, M, 3) - 0.5).astype(np.float32); rows = (rs.rand(B, N, 3) - 0.5).astype(np.float32)
k = min(N, M); rows[:, :k] = cols[:, :k] + 0.01 * rs.randn(B, k, 3).astype(np.float32)
for layout in ("pm", "cm"):
    for name, (form, norm, oform, onorm) in FORMS.items():
        if layout == "pm":
            tr, tc = torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda()
        else:
            tr = torch.from_numpy(np.ascontiguousarray(rows.transpose(0, 2, 1))).cuda().transpose(1, 2)
            tc = torch.from_numpy(np.ascontiguousarray(cols.transpose(0, 2, 1))).cuda().transpose(1, 2)
        r = F.nn1(tr, tc, form, norm, cache=False)
        o = O.nn1(oform, rows, cols, O.norms(onorm, rows), O.norms(onorm, cols))
        ra, ca = r.row_arg.cpu().numpy(), r.col_arg.cpu().numpy()
        bad_r = np.argwhere(ra != o.row_arg); bad_c = np.argwhere(ca != o.col_arg)
        print(layout, name, "row mismatches", len(bad_r), "col mismatches", len(bad_c), "values equal", np.array_equal(r.row_min.cpu().numpy(), o.row_min), np.array_equal(r.col_min.cpu().numpy(), o.col_min))
        for b, i in bad_r[:4]:
            print("    row", b, i, "ours", ra[b, i], "oracle", o.row_arg[b, i], "pos in group", o.row_arg[b, i] % 4)
        for b, j in bad_c[:4]:
            print("    col", b, j, "ours", ca[b, j], "oracle", o.col_arg[b, j], "pos in group", o.col_arg[b, j] % 4)
