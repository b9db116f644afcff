#!/usr/bin/env python
"""Generate tests/golden/hotpath_v1.npz from the oracles (run from the repo root).

The reference (QDax 0.5.1 on jax 0.8.0) cannot be imported in this image (no jax wheel, no
network), so these vectors come from oracle/qdx_oracle.c (exact-arithmetic spec) and
oracle/qdax_numpy.py (literal restatement); tests/test_oracle_golden.py re-derives them and
checks the two oracles against each other.  tools/dump_jax_golden.py writes the same keys
from the REAL reference when a JAX install is available; diffing the two files is the pin.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle as co  # noqa: E402
from oracle import jax_prng as jr  # noqa: E402
from oracle import qdax_numpy as qn  # noqa: E402


def main() -> None:
    out = {}
    # ---- PRNG
    out["split_key42"] = jr.split(jr.key(42))
    out["split_key0"] = jr.split(jr.key(0))
    out["split3_key7"] = jr.split(jr.key(7), 3)
    out["bits_key0_8"] = jr.random_bits(jr.key(0), (8,))
    out["uniform_key0_8"] = co.uniform(jr.key(0), 8)
    out["normal_key0_8"] = co.normal(jr.key(0), 8)
    out["normal_key42_1"] = co.normal(jr.key(42), 1)

    # ---- scenario S: 16x16 grid, D=20, B=64, half-filled repertoire
    rng = np.random.default_rng(20261017)
    K, D, B = 256, 20, 64
    cent = qn.compute_euclidean_centroids((16, 16), 0.0, 1.0)
    out["S_centroids"] = cent
    rep_f = np.where(rng.random(K) < 0.5, rng.standard_normal(K), -np.inf).astype(np.float32)
    rep_g = np.where(np.isinf(rep_f)[:, None], 0, rng.random((K, D))).astype(np.float32)
    rep_d = np.where(np.isinf(rep_f)[:, None], 0, cent).astype(np.float32)
    out["S_rep_g"], out["S_rep_f"], out["S_rep_d"] = rep_g, rep_f, rep_d
    key = jr.key(123)
    out["S_key"] = key
    out["S_select_idx"] = co.select_indices(rep_f, key, B)
    x, p1, p2 = co.emit_isoline(rep_g, rep_f, key, B, 0.05, 0.1, 0.0, 1.0)
    out["S_emit_x"], out["S_emit_p1"], out["S_emit_p2"] = x, p1, p2
    for task in ("arm", "rastrigin", "sphere"):
        f, d = co.score(task, x)
        out[f"S_{task}_f"], out[f"S_{task}_d"] = f, d
    f, d = co.score("arm", x)
    cells = co.cells(d, cent)
    out["S_arm_cells"] = cells
    # injected offspring with ties / NaN / -inf / -0.0 for the insert
    fi = np.round(rng.standard_normal(B), 1).astype(np.float32)
    fi[3], fi[7], fi[11], fi[12] = np.nan, -np.inf, -0.0, 0.0
    di = (rng.random((B, 2)) * 0.5).astype(np.float32)
    di[12] = di[11]
    ci = co.cells(di, cent)
    out["S_inj_f"], out["S_inj_d"], out["S_inj_cells"] = fi, di, ci
    for tb in ("first", "last"):
        g2, f2, d2, sidx = co.add(rep_g, rep_f, rep_d, x, fi, di, ci, tb)
        out[f"S_add_{tb}_g"], out[f"S_add_{tb}_f"], out[f"S_add_{tb}_d"], out[f"S_add_{tb}_sidx"] = g2, f2, d2, sidx

    # ---- mini C1: arm D=20, grid 16x16, B=64, 10 iterations from a 32-individual init
    init = co.uniform(jr.split(jr.key(42))[1], 32 * D).reshape(32, D)
    f0, d0 = co.score("arm", init)
    g0, ff0, dd0, _ = co.add(np.zeros((K, D)), np.full(K, -np.inf), np.zeros((K, 2)), init, f0, d0, co.cells(d0, cent))
    gN, fN, dN, kN, mN, _ = co.map_elites_scan(g0, ff0, dd0, cent, jr.key(5), 10, B, "arm")
    out["C1mini_init"] = init
    out["C1mini_g"], out["C1mini_f"], out["C1mini_d"], out["C1mini_key"], out["C1mini_metrics"] = gN, fN, dN, kN, mN

    # ---- DNS: N = 96 + 32, k = 3
    P, Bd = 96, 32
    pf = np.where(rng.random(P) < 0.8, np.round(rng.standard_normal(P), 1), -np.inf).astype(np.float32)
    pd = np.where(np.isinf(pf)[:, None], np.nan, rng.random((P, 2))).astype(np.float32)
    pg = rng.random((P, D)).astype(np.float32)
    bf = np.round(rng.standard_normal(Bd), 1).astype(np.float32)
    bd = rng.random((Bd, 2)).astype(np.float32)
    bg = rng.random((Bd, D)).astype(np.float32)
    g2, f2, d2, meta, surv = co.dns_add(pg, pf, pd, bg, bf, bd, 3)
    out.update(DNS_pg=pg, DNS_pf=pf, DNS_pd=pd, DNS_bg=bg, DNS_bf=bf, DNS_bd=bd, DNS_meta=meta, DNS_surv=surv,
               DNS_g=g2, DNS_f=f2, DNS_d=d2)

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "hotpath_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
