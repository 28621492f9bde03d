"""Fitting on the CUDA backbone: compass search with all 2P probes of a step in ONE launch, and the sweep of
(material x model x metric) fits that BASELINE configs[4] describes.

Mirrors include/optimizer/compass.h:82-140 of the reference (probe order +1, -1, +2, -2, ...; box test; the first
strictly better probe wins; contraction 0.5 when nothing improves; converged when step < tolerance) on top of
bbmcu_loss_eval's K-parameter-set batches.  Host logic only - every loss value comes from the loss kernels.
"""
import numpy as np

from . import Bsdf, METRICS, ATTR_ALL

EPS32 = float(np.finfo(np.float32).eps)


class CompassBatched:
    """optimization_algorithm (include/concepts/optimization_algorithm.h:24-37): step() -> loss, reset(), is_converged()"""

    def __init__(self, loss, bsdf, lower=None, upper=None, tolerance=EPS32, step_size=1.0, contraction=0.5, expansion=1.0):
        self.loss, self.bsdf = loss, bsdf
        self.param = np.asarray(bsdf.parameter_values(), np.float32).astype(np.float64)
        self.lower = None if lower is None else np.asarray(lower, np.float32)
        self.upper = None if upper is None else np.asarray(upper, np.float32)
        self.tolerance, self.initial_step = np.float32(tolerance), np.float32(step_size)
        self.contraction, self.expansion = np.float32(contraction), np.float32(expansion)
        self.reset()

    def reset(self):
        self.step_size = self.initial_step
        self.loss_value = np.float32(self.loss(self.bsdf, self.param[None])[0])

    def is_converged(self):
        return bool(self.step_size < self.tolerance)

    def probes(self):
        """(2P, P) probe vectors in the reference's order and the in-box mask"""
        P = len(self.param)
        base = self.param.astype(np.float32)
        pr = np.repeat(base[None], 2 * P, 0)
        j = np.arange(2 * P) // 2
        sgn = np.where(np.arange(2 * P) % 2 == 0, np.float32(1), np.float32(-1))
        pr[np.arange(2 * P), j] = base[j] + sgn * self.step_size          # float addition, like Value + _step
        ok = np.ones(2 * P, bool)
        # compass.h:112-126 tests ALL parameters of the probed vector against the box, not only the probed coordinate: a
        # start point outside the box in coordinate i rejects every probe of the other coordinates
        inside = np.ones(P, bool)
        if self.lower is not None:
            ok &= pr[np.arange(2 * P), j] >= self.lower[j]
            inside &= base >= self.lower
        if self.upper is not None:
            ok &= pr[np.arange(2 * P), j] <= self.upper[j]
            inside &= base <= self.upper
        others_inside = np.array([inside[np.arange(P) != jj].all() for jj in range(P)])
        ok &= others_inside[j]
        pr[~ok] = base                                                     # evaluated but ignored: fixed launch shape
        return pr.astype(np.float64), ok

    def step(self):
        if self.is_converged():
            return 0.0
        pr, ok = self.probes()
        err = self.loss(self.bsdf, pr).astype(np.float32)                  # ONE launch
        err = np.where(ok & np.isfinite(err), err, np.float32(np.inf))
        k = int(np.argmin(err))                                            # first minimum = first strictly better probe
        if err[k] < self.loss_value:
            self.param = pr[k].copy()
            self.loss_value = err[k]
            self.step_size = np.float32(self.expansion * self.step_size)
        else:
            self.step_size = np.float32(self.contraction * self.step_size)
        return float(self.loss_value)

    def commit(self):
        self.bsdf.set_parameter_values(self.param)
        return self.bsdf


class CompassMulti:
    """M independent compass searches - one per material of a batched loss (Loss over a list of measured tables) - advanced
    together: all M x 2P probes of a step go through ONE launch (Loss.eval_multi).  Per material the search is exactly
    CompassBatched's (same probes, same acceptance rule); materials that have converged keep their point."""

    def __init__(self, loss, bsdf, lower=None, upper=None, tolerance=EPS32, step_size=1.0, contraction=0.5, expansion=1.0, start=None):
        self.loss, self.bsdf = loss, bsdf
        self.M = loss.materials()
        p0 = np.asarray(bsdf.parameter_values(), np.float32).astype(np.float64)
        self.param = np.tile(p0, (self.M, 1)) if start is None else np.asarray(start, np.float32).astype(np.float64).reshape(self.M, -1)
        self.P = self.param.shape[1]
        self.lower = None if lower is None else np.asarray(lower, np.float32)
        self.upper = None if upper is None else np.asarray(upper, np.float32)
        self.tolerance, self.initial_step = np.float32(tolerance), np.float32(step_size)
        self.contraction, self.expansion = np.float32(contraction), np.float32(expansion)
        self.reset()

    def reset(self):
        self.step_size = np.full(self.M, self.initial_step, np.float32)
        self.loss_value = self.loss.eval_multi(self.bsdf, self.param[:, None, :])[:, 0].astype(np.float32)
        self.steps = 0

    def is_converged(self):
        return self.step_size < self.tolerance

    def step(self):
        """one step of every unconverged search; returns the (M,) loss values"""
        M, P = self.M, self.P
        k = np.arange(2 * P)
        j = k // 2
        sgn = np.where(k % 2 == 0, np.float32(1), np.float32(-1))
        base = self.param.astype(np.float32)                                   # (M, P)
        pr = np.repeat(base[:, None, :], 2 * P, 1)                             # (M, 2P, P)
        moved = base[:, j] + sgn[None, :] * self.step_size[:, None]            # float addition, like Value + _step
        pr[:, k, j] = moved
        ok = np.ones((M, 2 * P), bool)
        inside = np.ones((M, P), bool)
        if self.lower is not None:
            ok &= moved >= self.lower[j][None]
            inside &= base >= self.lower[None]
        if self.upper is not None:
            ok &= moved <= self.upper[j][None]
            inside &= base <= self.upper[None]
        n_out = (~inside).sum(1)                                               # all OTHER coordinates inside the box (compass.h:112-126)
        ok &= (n_out[:, None] - (~inside)[:, j]) == 0
        active = ~self.is_converged()
        ok &= active[:, None]
        pr[~ok] = np.repeat(base[:, None, :], 2 * P, 1)[~ok]                   # evaluated but ignored: fixed launch shape
        err = self.loss.eval_multi(self.bsdf, pr.astype(np.float64)).astype(np.float32)     # ONE launch
        err = np.where(ok & np.isfinite(err), err, np.float32(np.inf))
        best = np.argmin(err, 1)                                               # first minimum = first strictly better probe
        e_best = err[np.arange(M), best]
        better = active & (e_best < self.loss_value)
        self.param[better] = pr[np.arange(M), best][better].astype(np.float64)
        self.loss_value = np.where(better, e_best, self.loss_value)
        self.step_size = np.where(active, np.where(better, self.expansion * self.step_size, self.contraction * self.step_size), self.step_size).astype(np.float32)
        self.steps += 1
        return self.loss_value


def fit(ctx, fitted_string, reference, metric="nganL2", grid=None, max_steps=200, box=True):
    """fit `fitted_string` (a BSDF string; its values are the start) to `reference` (Bsdf or MERL table)"""
    b = Bsdf(fitted_string)
    L = ctx.loss(metric, reference, grid)
    opt = CompassBatched(L, b, b.parameter_lower_bound() if box else None, b.parameter_upper_bound() if box else None)
    trace = []
    for _ in range(max_steps):
        if opt.is_converged():
            break
        trace.append(opt.step())
    opt.commit()
    return b, np.array(trace)


# relative cost of one loss pass per model family, used to balance a sweep over ranks (the He
# family dominates); measured single-model eval rates of BASELINE.md section 2, inverted and rounded
MODEL_COST = {"He": 40.0, "HeWestin": 40.0, "HeHolzschuch": 15.0, "NganHe": 40.0, "Bagher": 6.0, "EPD": 5.0,
              "Ribardiere": 4.0, "RibardiereAnisotropic": 4.0}


def sweep_jobs(materials, models, metrics=METRICS):
    """the (material, model, metric) job list of configs[4] and a cost estimate per job"""
    jobs = [(m, mod, met) for m in materials for mod in models for met in metrics]
    cost = [MODEL_COST.get(mod, 1.0) for _, mod, _ in jobs]
    return jobs, cost


def fitted_string_for(model):
    """every .fit entry of the reference is Aggregate(Lambertian, X); models with their own diffuse term stand alone"""
    if model in ("Lambertian", "OrenNayar", "AshikhminShirleyFull"):
        return f"{model}()"
    return f"Aggregate(Lambertian(), {model}())"


def run_sweep_by_material(ctx, tables, models, metrics=METRICS, rank=0, world=1, max_steps=50, grid=None, loss=None, progress=None):
    """BASELINE configs[4] the way the hardware wants it: the sweep is split BY MATERIAL; a rank keeps
    its materials' measured tables resident as ONE batched loss object (17.5 MB each, uploaded once) and, per (model, metric),
    advances the compass searches of all its materials together - one launch per step for M x 2P probes.  The metric is a
    per-call switch, so the six metrics share the tabulated data.  No collective.
    `tables`: {material: (3, 1458000) float32}; returns {(material, model, metric): (bsdf string, final loss, steps)}."""
    import time
    from .shard import shard_range
    names = sorted(tables)
    first, count = shard_range(len(names), rank, world)
    mine = names[first:first + count]
    out = {}
    if not mine:
        return out
    L = loss if loss is not None else ctx.loss(metrics[0], [tables[m] for m in mine], grid)
    for mod in models:
        b = Bsdf(fitted_string_for(mod))
        lo, hi = b.parameter_lower_bound(), b.parameter_upper_bound()
        for met in metrics:
            t0 = time.perf_counter()
            L.set_metric(met)
            opt = CompassMulti(L, b, lo, hi)
            for _ in range(max_steps):
                if opt.is_converged().all():
                    break
                opt.step()
            for m, mat in enumerate(mine):
                fb = Bsdf(fitted_string_for(mod))
                fb.set_parameter_values(opt.param[m])
                out[(mat, mod, met)] = (fb.to_string(), float(opt.loss_value[m]), opt.steps)
            if progress is not None:
                progress(mod, met, time.perf_counter() - t0, opt.steps, len(mine), 2 * opt.P)
    return out


def run_sweep(ctx, tables, models, metrics=METRICS, rank=0, world=1, max_steps=50, grid=None):
    """this rank's share of the sweep.  `tables`: {material: (3, 1458000) float32 MERL table}.  Returns
    {(material, model, metric): (bsdf string, final loss, steps)}; no collective is involved."""
    from .shard import partition_by_cost
    jobs, cost = sweep_jobs(sorted(tables), models, metrics)
    mine = partition_by_cost(cost, world)[rank]
    out = {}
    for j in mine:
        mat, mod, met = jobs[j]
        b, trace = fit(ctx, fitted_string_for(mod), tables[mat], met, grid, max_steps)
        out[jobs[j]] = (b.to_string(), float(trace[-1]) if len(trace) else float("nan"), len(trace))
    return out
