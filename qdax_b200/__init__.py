"""qdax_b200 -- B200-native MAP-Elites generation step behind QDax's Python surface.

Import paths mirror the reference (`qdax.core.map_elites` -> `qdax_b200.core.map_elites`, ...), so the
README example of QDax runs with the import root changed and `jax.random` -> `qdax_b200.random`,
`jax.lax.scan` -> `qdax_b200.lax.scan`.  Everything computes on CUDA through libqdx.so (include/qdx.h);
there is no CPU fallback.
"""

__version__ = "0.1.0"
