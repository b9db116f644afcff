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
    np.savez_compressed(out, **o)
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "jax_golden.npz")
