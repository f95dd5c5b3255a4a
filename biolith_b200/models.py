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

# ---- keyword surface of the reference models (occu.py:19-40, occu_rn.py:20-40, occu_cop.py:18-41, nmixture.py:19-38,
# occu_cs.py:18-40).  Every key a caller may pass is listed; anything else is an error, and every listed key is
# either honoured by the kernels or rejected when it asks for something outside the accelerated path.
_COMMON = {"coords", "ell", "n_species", "prior_beta", "prior_alpha", "regressor_det", "prior_gp_sd", "prior_gp_length",
           "site_random_effects", "obs_random_effects", "prior_site_re_sd", "prior_obs_re_sd"}
KEYWORDS = {
    "occu": _COMMON | {"false_positives_constant", "false_positives_unoccupied", "regressor_occ",
                       "prior_prob_fp_constant", "prior_prob_fp_unoccupied"},
    "occu_rn": _COMMON | {"false_positives_constant", "max_abundance", "regressor_abu", "prior_prob_fp_constant"},
    "occu_cop": _COMMON | {"false_positives_constant", "false_positives_unoccupied", "regressor_occ",
                           "prior_rate_fp_constant", "prior_rate_fp_unoccupied"},
    "nmixture": _COMMON | {"max_abundance", "regressor_abu"},
    "occu_cs": _COMMON | {"regressor_occ", "prior_mu", "prior_sigma"},
}
# ell / prior_gp_* / prior_*_re_sd belong to parts that are switched off here (coords, random effects are rejected
# when on), so like in the reference they have no effect and need no mapping


def _dist_params(pr, kind, fields):
    """Duck-typed read of a numpyro distribution (no numpyro import): class name + float parameters."""
    if type(pr).__name__ != kind:
        return None
    try:
        vals = tuple(float(getattr(pr, f)) for f in fields)
    except (AttributeError, TypeError, ValueError):
        return None
    return vals


def model_options(name, kwargs):
    """Validate the model keywords handed to ``fit`` -> (fp_constant, fp_unoccupied, max_abundance, prior_kw).

    Raises BiolithB200Error for unknown keys, for options outside the accelerated path and for priors the
    kernels cannot carry -- never a silent default (the reference would sample under the given prior)."""
    from ._lib import BiolithB200Error

    def bad(msg, code=-2):
        return BiolithB200Error(code, "fit", msg)

    re_kw = {}
    unknown = set(kwargs) - KEYWORDS[name]
    if unknown:
        raise bad(f"unknown keyword(s) for {name}: {sorted(unknown)} (accepted: {sorted(KEYWORDS[name])})", -1)
    if kwargs.get("coords") is not None:
        raise bad("coords (spatial HSGP effect) is outside the accelerated path (no fallback)")
    for k, pk, out in (("site_random_effects", "prior_site_re_sd", "prior_site_re_sd_scale"),
                       ("obs_random_effects", "prior_obs_re_sd", "prior_obs_re_sd_scale")):
        if not kwargs.get(k):
            continue
        if name != "occu" or kwargs.get("false_positives_constant") or kwargs.get("false_positives_unoccupied"):
            raise bad(f"{k} is accelerated for occu without false-positive extras only (no fallback)")
        re_kw[k] = True
        pr = kwargs.get(pk)
        if pr is not None:  # occu.py:38-39: HalfNormal(scale)
            v = _dist_params(pr, "HalfNormal", ("scale",))
            if v is None:
                raise bad(f"{pk}: only HalfNormal(scale) priors are accelerated")
            re_kw[out] = v[0]
    for k in ("regressor_occ", "regressor_det", "regressor_abu"):
        r = kwargs.get(k)
        if r is not None and getattr(r, "__name__", "") != "LinearRegression":
            raise bad(f"{k}={r!r}: only LinearRegression is accelerated")
    fpc = bool(kwargs.get("false_positives_constant", False))
    fpu = bool(kwargs.get("false_positives_unoccupied", False))
    if fpc and fpu:  # occu.py:127-129, occu_cop.py:113-115 assert the same
        raise bad("false_positives_constant and false_positives_unoccupied cannot both be True", -1)
    prior_kw = {}
    for k in ("prior_beta", "prior_alpha"):
        pr = kwargs.get(k)
        if pr is not None:
            v = _dist_params(pr, "Normal", ("loc", "scale"))
            if v is None:
                raise bad(f"{k}: only Normal(loc, scale) priors are accelerated")
            prior_kw[k] = v
    # false-positive priors: Beta(a, b) on the probability, Exponential(rate) on the rate; only the prior of the
    # ENABLED parameter matters (the flags are exclusive), a prior for a disabled one is inert as in the reference
    for flag, key in ((fpc, "prior_prob_fp_constant"), (fpu, "prior_prob_fp_unoccupied")):
        pr = kwargs.get(key)
        if pr is not None and flag:
            v = _dist_params(pr, "Beta", ("concentration1", "concentration0"))
            if v is None:
                raise bad(f"{key}: only Beta(a, b) priors are accelerated")
            prior_kw["prior_fp_beta"] = v
    for flag, key in ((fpc, "prior_rate_fp_constant"), (fpu, "prior_rate_fp_unoccupied")):
        pr = kwargs.get(key)
        if pr is not None and flag:
            v = _dist_params(pr, "Exponential", ("rate",))
            if v is None:
                raise bad(f"{key}: only Exponential(rate) priors are accelerated")
            prior_kw["prior_fp_rate"] = v[0]
    if name == "occu_cs":
        # occu_cs.py:29-30: one distribution or a (f = 0, f = 1) pair; the kernels carry one Normal(0, s) / Gamma(a, b)
        pr = kwargs.get("prior_mu")
        if pr is not None:
            v = None if isinstance(pr, tuple) else _dist_params(pr, "Normal", ("loc", "scale"))
            if v is None or v[0] != 0.0:
                raise bad("prior_mu: only a single zero-centred Normal prior is accelerated")
            prior_kw["prior_mu_scale"] = v[1]
        pr = kwargs.get("prior_sigma")
        if pr is not None:
            v = None if isinstance(pr, tuple) else _dist_params(pr, "Gamma", ("concentration", "rate"))
            if v is None:
                raise bad("prior_sigma: only a single Gamma(concentration, rate) prior is accelerated")
            prior_kw["prior_sigma"] = v
    prior_kw.update(re_kw)
    return fpc, fpu, int(kwargs.get("max_abundance", 100)), prior_kw
