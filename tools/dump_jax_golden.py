#!/usr/bin/env python
"""Regenerate the golden vectors of tests/golden/hotpath_v1.npz from the REAL reference (QDax 0.5.1 on jax 0.8.0).

Cannot run in this image (no jax wheel, no network).  On a machine with `pip install qdax==0.5.1 jax==0.8.0`:

    python tools/dump_jax_golden.py out.npz && python - <<'PY'
    import numpy as np; a, b = np.load("out.npz"), np.load("tests/golden/hotpath_v1.npz")
    for k in a.files: print(k, np.array_equal(a[k], b[k], equal_nan=True), float(np.nanmax(np.abs(a[k].astype(float) - b[k].astype(float)))) if a[k].dtype.kind == "f" else "")
    PY

Integer arrays (split keys, bits, select indices, cells, scatter indices) must match exactly; float arrays are
expected to agree to ~1e-6 (XLA's own log1p / sin / cos and reduction orders differ in the last bits).
Any exact mismatch in an integer array pins a wrong assumption of the oracle and must be fixed there first.
"""
import functools
import sys

import numpy as np


def main(out):
    import jax
    import jax.numpy as jnp
    from qdax.core.containers.mapelites_repertoire import MapElitesRepertoire, compute_euclidean_centroids, get_cells_indices
    from qdax.core.emitters.mutation_operators import isoline_variation
    from qdax.core.emitters.standard_emitters import MixingEmitter
    from qdax.tasks.arm import arm_scoring_function
    from qdax.tasks.standard_functions import rastrigin_scoring_function, sphere_scoring_function

    ref = np.load("tests/golden/hotpath_v1.npz")
    kd = lambda k: np.asarray(jax.random.key_data(k))
    wrap = lambda w: jax.random.wrap_key_data(jnp.asarray(w, dtype=jnp.uint32))
    o = {}
    o["split_key42"] = kd(jax.random.split(jax.random.key(42)))
    o["split_key0"] = kd(jax.random.split(jax.random.key(0)))
    o["split3_key7"] = kd(jax.random.split(jax.random.key(7), 3))
    o["bits_key0_8"] = np.asarray(jax.random.bits(jax.random.key(0), (8,), dtype=jnp.uint32))
    o["uniform_key0_8"] = np.asarray(jax.random.uniform(jax.random.key(0), (8,)))
    o["normal_key0_8"] = np.asarray(jax.random.normal(jax.random.key(0), (8,)))
    o["normal_key42_1"] = np.asarray(jax.random.normal(jax.random.key(42), (1,)))
    cent = compute_euclidean_centroids((16, 16), 0.0, 1.0)
    o["S_centroids"] = np.asarray(cent)
    rep = MapElitesRepertoire(genotypes=jnp.asarray(ref["S_rep_g"]), fitnesses=jnp.asarray(ref["S_rep_f"]).reshape(-1, 1),
                              descriptors=jnp.asarray(ref["S_rep_d"]), centroids=cent, extra_scores={}, keys_extra_scores=())
    key = wrap(ref["S_key"])
    B = ref["S_emit_x"].shape[0]
    sel = rep.select(key, B)
    o["S_select_genotypes"] = np.asarray(sel.genotypes)          # compare with S_rep_g[S_select_idx]
    em = MixingEmitter(lambda x, k: x, functools.partial(isoline_variation, iso_sigma=0.05, line_sigma=0.1, minval=0.0, maxval=1.0), 1.0, B)
    x, _ = em.emit(rep, None, key)
    o["S_emit_x"] = np.asarray(x)
    for name, fn in (("arm", arm_scoring_function), ("rastrigin", rastrigin_scoring_function), ("sphere", sphere_scoring_function)):
        f, d, _ = fn(jnp.asarray(ref["S_emit_x"]), key)
        o[f"S_{name}_f"], o[f"S_{name}_d"] = np.asarray(f), np.asarray(d)
    o["S_arm_cells"] = np.asarray(get_cells_indices(jnp.asarray(ref["S_arm_d"]), cent))
    o["S_inj_cells"] = np.asarray(get_cells_indices(jnp.asarray(ref["S_inj_d"]), cent))
    new = rep.add(jnp.asarray(ref["S_emit_x"]), jnp.asarray(ref["S_inj_d"]), jnp.asarray(ref["S_inj_f"]))
    o["S_add_jax_g"], o["S_add_jax_f"], o["S_add_jax_d"] = np.asarray(new.genotypes), np.asarray(new.fitnesses).ravel(), np.asarray(new.descriptors)
    # ---- rows added after round 1: pytree isoline, MELS, DNS -- inputs are generated here from fixed NumPy seeds, compare with
    #      oracle/qdax_numpy.py (isoline_variation_tree, mels_add, dns_add) on the same inputs
    from qdax.core.containers.dns_repertoire import DominatedNoveltyRepertoire
    from qdax.core.containers.mels_repertoire import MELSRepertoire

    rng = np.random.default_rng(0)
    shapes = [(2, 3), (5,), (3, 2, 2)]
    t1 = {f"leaf{i}": jnp.asarray(rng.random((23,) + s).astype(np.float32)) for i, s in enumerate(shapes)}
    t2 = {f"leaf{i}": jnp.asarray(rng.random((23,) + s).astype(np.float32)) for i, s in enumerate(shapes)}
    tv = isoline_variation(t1, t2, jax.random.key(3), iso_sigma=0.05, line_sigma=0.1, minval=0.0, maxval=1.0)
    for i in range(len(shapes)):
        o[f"T_x1_{i}"], o[f"T_x2_{i}"], o[f"T_out_{i}"] = np.asarray(t1[f"leaf{i}"]), np.asarray(t2[f"leaf{i}"]), np.asarray(tv[f"leaf{i}"])
    cent6 = compute_euclidean_centroids((6, 6), 0.0, 1.0)
    mels = MELSRepertoire.init_default(genotype=jnp.zeros(8), centroids=cent6)
    mg = rng.random((64, 8)).astype(np.float32)
    md = (rng.random((64, 1, 2)) + 0.15 * rng.standard_normal((64, 5, 2))).astype(np.float32)
    mf = rng.standard_normal((64, 5)).astype(np.float32)
    mels = mels.add(jnp.asarray(mg), jnp.asarray(md), jnp.asarray(mf))
    o["M_g"], o["M_d"], o["M_f"] = mg, md, mf
    o["M_out_g"], o["M_out_f"], o["M_out_d"], o["M_out_s"] = (np.asarray(mels.genotypes), np.asarray(mels.fitnesses).ravel(),
                                                             np.asarray(mels.descriptors), np.asarray(mels.spreads))
    dns = DominatedNoveltyRepertoire.init(genotypes=jnp.asarray(ref["DNS_pg"]), fitnesses=jnp.asarray(ref["DNS_pf"]).reshape(-1, 1),
                                          descriptors=jnp.asarray(ref["DNS_pd"]), population_size=ref["DNS_pg"].shape[0], k=3)
    dns = dns.add(jnp.asarray(ref["DNS_bg"]), jnp.asarray(ref["DNS_bd"]), jnp.asarray(ref["DNS_bf"]).reshape(-1, 1))
    o["DNS_jax_g"], o["DNS_jax_f"], o["DNS_jax_d"] = np.asarray(dns.genotypes), np.asarray(dns.fitnesses).ravel(), np.asarray(dns.descriptors)
    np.savez_compressed(out, **o)
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "jax_golden.npz")
