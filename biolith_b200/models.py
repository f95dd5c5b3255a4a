"""Model identifiers mirroring ``biolith.models`` for the three likelihoods on the accelerated path.

In the reference these are NumPyro model functions (biolith/models/occu.py:19, occu_rn.py:20,
occu_cop.py:18) handed to ``fit``.  Here they are lightweight descriptors with the same names and
keyword surface; ``biolith_b200.fit`` also accepts the reference's own functions (matched by
``__name__``), so ``fit(biolith.models.occu, **data)`` and ``fit(biolith_b200.models.occu, **data)``
select the same kernels.  Options outside the accelerated path raise (no silent fallback).
"""

from __future__ import annotations


class _Model:
    def __init__(self, name, doc):
        self.__name__ = name
        self.__doc__ = doc

    def __call__(self, *a, **k):
        raise RuntimeError(
            f"biolith_b200.models.{self.__name__} is a descriptor for biolith_b200.fit(); the NumPyro program "
            f"lives in biolith.models.{self.__name__}")

    def __repr__(self):
        return f"<biolith_b200 model {self.__name__}>"


occu = _Model("occu", "Bernoulli occupancy model (MacKenzie et al. 2002); biolith/models/occu.py:19-242")
occu_rn = _Model("occu_rn", "Royle-Nichols abundance-induced heterogeneity; biolith/models/occu_rn.py:20-222")
occu_cop = _Model("occu_cop", "Count-detection occupancy (Pautrel et al. 2024); biolith/models/occu_cop.py:18-255")

nmixture = _Model("nmixture", "N-mixture model for repeated counts (Royle 2004); biolith/models/nmixture.py:19-220")

occu_cs = _Model("occu_cs", "Continuous-score occupancy (Rhinehart et al. 2022); biolith/models/occu_cs.py:18-223")

SUPPORTED = {"occu": occu, "occu_rn": occu_rn, "occu_cop": occu_cop, "nmixture": nmixture, "occu_cs": occu_cs}

# keyword arguments of the reference models that the accelerated path honours / must reject
HONOURED = {"false_positives_constant", "false_positives_unoccupied", "max_abundance", "n_species"}
REJECTED_IF_SET = {
    "coords": None, "site_random_effects": False, "obs_random_effects": False,
}
