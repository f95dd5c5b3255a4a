"""Host-side ingestion with the behaviour of ``biolith.utils.data`` (biolith/utils/data.py:9-165).

``prepare_data`` accepts what the reference's ``fit`` accepts -- numpy arrays or pandas DataFrames -- and
returns arrays in the model layout plus the covariate names used by ``rename_samples``:

* DataFrames are re-ordered to ONE common row order: the index of the first DataFrame among
  ``obs, site_covs, obs_covs, session_duration`` (data.py:14-40); a frame is only re-ordered when its index
  holds exactly the same labels;
* ``obs_covs`` frames must carry MultiIndex columns (covariate, replicate) or (covariate, period, replicate)
  and are reshaped to ``(S, P, J, Ko)`` (data.py:45-77); covariate names come from ``levels[0]``;
* ``obs`` / ``session_duration`` frames with (period, replicate) MultiIndex columns become ``(S, P, J)``
  (data.py:78-109);
* lower-rank arrays get the period dimension inserted (data.py:113-127);
* names default to "0", "1", ... (data.py:129-132).

Runs once per fit on the host; nothing here is on the per-leapfrog path.
"""

from __future__ import annotations

import numpy as np


def _is_frame(x) -> bool:
    return hasattr(x, "index") and hasattr(x, "columns") and hasattr(x, "to_numpy")


def _common_order(frames):
    for f in frames:
        if _is_frame(f):
            return f.index
    return None


def _reorder(df, order):
    if order is None or not _is_frame(df):
        return df
    have = df.index
    same_labels = len(have) == len(order) and order.isin(have).all() and have.isin(order).all()
    if same_labels and not have.equals(order):
        return df.loc[order]
    return df


def _levels(columns):
    return [len(lv) for lv in columns.levels]


def prepare_data(site_covs=None, obs_covs=None, obs=None, session_duration=None):
    """-> (site_covs, obs_covs, obs, session_duration, site_covs_names, obs_covs_names) as numpy arrays."""
    order = _common_order((obs, site_covs, obs_covs, session_duration))
    site_covs, obs_covs, obs, session_duration = (
        _reorder(a, order) for a in (site_covs, obs_covs, obs, session_duration))
    site_names = obs_names = None

    if _is_frame(site_covs):
        site_names = ["intercept"] + list(site_covs.columns)
        site_covs = site_covs.to_numpy()

    if _is_frame(obs_covs):
        cols = obs_covs.columns
        if getattr(cols, "nlevels", 1) < 2:
            raise ValueError("obs_covs DataFrame must use MultiIndex columns with levels "
                             "(covariate, period, replicate) for multi-season data.")
        n = _levels(cols)
        obs_names = ["intercept"] + list(cols.levels[0])
        flat = obs_covs.to_numpy()
        if len(n) == 2:    # (covariate, replicate) -> (S, 1, J, Ko)
            obs_covs = np.moveaxis(flat.reshape(flat.shape[0], n[0], n[1]), 1, 2)[:, None]
        elif len(n) == 3:  # (covariate, period, replicate) -> (S, P, J, Ko)
            obs_covs = np.moveaxis(flat.reshape(flat.shape[0], n[0], n[1], n[2]), 1, 3)
        else:
            raise ValueError("obs_covs with MultiIndex columns must have 2 or 3 levels.")

    def _period_replicate_frame(df, what):
        if not _is_frame(df):
            return df
        cols = df.columns
        if getattr(cols, "nlevels", 1) == 1:
            return df.to_numpy()
        n = _levels(cols)
        if len(n) != 2:
            raise ValueError(f"{what} with MultiIndex columns must have 2 levels.")
        return df.to_numpy().reshape(df.shape[0], n[0], n[1])

    session_duration = _period_replicate_frame(session_duration, "session_duration")
    obs = _period_replicate_frame(obs, "obs")

    def _with_period_dim(a, name):
        if a is None:
            return None
        a = np.asarray(a)
        if name == "obs_covs":
            if a.ndim == 2:
                a = a[:, :, None]
            if a.ndim == 3:
                a = a[:, None, :, :]
        elif a.ndim == 2:
            a = a[:, None, :]
        return a

    obs_covs = _with_period_dim(obs_covs, "obs_covs")
    obs = _with_period_dim(obs, "obs")
    session_duration = _with_period_dim(session_duration, "session_duration")
    site_covs = None if site_covs is None else np.asarray(site_covs)

    if site_names is None and site_covs is not None:
        site_names = ["0"] + [str(i + 1) for i in range(site_covs.shape[1])]
    if obs_names is None and obs_covs is not None:
        obs_names = ["0"] + [str(i + 1) for i in range(obs_covs.shape[-1])]
    return site_covs, obs_covs, obs, session_duration, site_names, obs_names


def rename_samples(samples, site_covs_names=None, obs_covs_names=None):
    """``beta`` -> ``cov_state_<name>``, ``alpha`` -> ``cov_det_<name>`` (biolith/utils/data.py:145-165)."""
    samples = dict(samples)
    for key, prefix, names in (("beta", "cov_state_", site_covs_names), ("alpha", "cov_det_", obs_covs_names)):
        if names is None:
            continue
        for i, n in enumerate(names):  # regressors that register one site per coefficient
            if f"{key}_{i}" in samples:
                samples[f"{prefix}{n}"] = samples.pop(f"{key}_{i}")
        if key in samples:
            coef = samples.pop(key)
            for i, n in enumerate(names):
                samples[f"{prefix}{n}"] = coef[..., i]
    return samples
