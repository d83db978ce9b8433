"""Geometry optimisers: drop-in for ``micmec.sampling.opt`` (opt.py:52-443).

``QNOptimizer`` is the trust-radius quasi-Newton scheme the reference's scripts use (``optimisation.py``,
``stress_strain.py``): an SR1 (or BFGS) model of the Hessian, a step from the model's spectrum restricted to the trust
radius, acceptance on an energy decrease.  Every ``fun`` call is one force evaluation on the GPU through the DOF object.
``CGOptimizer`` depends on ``molmod.minimizer`` in the reference; it is only available when that package is.
Ensembles of small systems are optimised in lockstep by ``micmec_b200.sampling.batchopt.ReplicaQNOptimizer``.
"""
import time

import numpy as np

from .iterative import (
    Iterative, AttributeStateItem, PosStateItem, VolumeStateItem, DomainStateItem, EPotContribStateItem, Hook,
)

__all__ = [
    "OptScreenLog", "BaseOptimizer", "CGOptimizer", "HessianModel", "BFGSHessianModel", "SR1HessianModel", "QNOptimizer",
    "solve_trust_radius",
]


class OptScreenLog(Hook):
    """Convergence table on the screen (opt.py:52-84).  Prints when ``verbose`` is True or, by default, when the
    package log level is at least medium (see ``VerletScreenLog``)."""

    def __init__(self, start=0, step=1, verbose=None):
        Hook.__init__(self, start, step)
        self.time0 = None
        self._verbose = verbose
        self.lines = 0

    @property
    def verbose(self):
        if self._verbose is None:
            from ..log import log

            return log.do_medium
        return self._verbose

    def __call__(self, iterative):
        if self.time0 is None:
            self.time0 = time.time()
            if self.verbose:
                print("counter  Conv.val.  N           Worst         Energy   Walltime")
        self.lines += 1
        if self.verbose:
            dof = iterative.dof
            print("%7i % 10.3e %2i %15s % 14.8e %10.1f" % (
                iterative.counter, dof.conv_val, dof.conv_count, dof.conv_worst, iterative.epot, time.time() - self.time0))


class BaseOptimizer(Iterative):
    default_state = [
        AttributeStateItem("counter"), AttributeStateItem("epot"), PosStateItem(), VolumeStateItem(), DomainStateItem(),
        EPotContribStateItem(),
    ]
    log_name = "XXOPT"

    def __init__(self, dof, state=None, hooks=None, counter0=0):
        self.dof = dof
        Iterative.__init__(self, dof.mmf, state, hooks, counter0)

    def _add_default_hooks(self):
        if not any(isinstance(hook, OptScreenLog) for hook in self.hooks):
            self.hooks.append(OptScreenLog())

    def fun(self, x, do_gradient=False):
        if do_gradient:
            self.epot, gx = self.dof.fun(x, True)
            return self.epot, gx
        self.epot = self.dof.fun(x, False)
        return self.epot

    def initialize(self):
        # never flags convergence the first time; it records the reference geometry for the displacement criteria
        self.dof.check_convergence()
        Iterative.initialize(self)

    def propagate(self):
        self.dof.check_convergence()
        Iterative.propagate(self)
        return self.dof.converged

    def finalize(self):
        self.dof.log()


class _CGMinimizer(object):
    """Self-contained nonlinear conjugate gradients with a Newton line search: the algorithm family the reference
    borrows from ``molmod.minimizer`` (``Minimizer(x0, fun, ConjugateGradient(), NewtonLineSearch(), ...)``,
    opt.py:160-164).  molmod is not vendored with the reference, so this is a restatement of the published scheme, not
    a line-by-line mirror (parity with molmod's iterates is unpinned; the minimum it converges to is tested):

    * search direction: Polak-Ribiere with automatic restart, ``beta = max(0, g.(g - g_old) / g_old.g_old)``;
      the first step and every restart are steepest descent;
    * line search: one Newton step along the unit direction, with the curvature from a finite difference of the
      ANALYTIC gradient, ``f'' ~ (g(x + eps d) - g(x)).d / eps``; the step is limited to ``qmax`` and halved until the
      function decreases; a non-positive curvature falls back to a step of length ``qmax``;
    * a failed line search along a conjugate direction restarts with steepest descent; a failed steepest-descent line
      search ends the minimisation (``propagate`` returns False), as in the reference's wiring.
    """

    def __init__(self, x0, fun, eps=1e-6, qmax=1.0, max_halvings=20):
        self.x = np.array(x0, dtype=float)
        self.fun = fun
        self.eps, self.qmax, self.max_halvings = eps, qmax, max_halvings
        self.f = self.gradient = self.direction = self.gradient_old = None
        self.status = "SD"

    def initialize(self):
        self.f, self.gradient = self.fun(self.x, do_gradient=True)
        self.gradient = np.array(self.gradient, dtype=float)
        self.direction = -self.gradient
        self.status = "SD"

    def _line_search(self):
        norm = np.linalg.norm(self.direction)
        if norm == 0.0:
            return False
        unit = self.direction / norm
        slope = np.dot(self.gradient, unit)
        if slope >= 0.0:
            return False  # not a descent direction
        _, g_eps = self.fun(self.x + self.eps * unit, do_gradient=True)
        curvature = (np.dot(g_eps, unit) - slope) / self.eps
        step = -slope / curvature if curvature > 0.0 else self.qmax
        step = min(step, self.qmax)
        for _ in range(self.max_halvings):
            xnew = self.x + step * unit
            fnew, gnew = self.fun(xnew, do_gradient=True)
            if np.isfinite(fnew) and fnew < self.f:
                self.x, self.f = xnew, fnew
                self.gradient_old, self.gradient = self.gradient, np.array(gnew, dtype=float)
                return True
            step *= 0.5
        return False

    def propagate(self):
        if not self._line_search():
            if self.status == "SD":
                return False
            self.direction, self.status = -self.gradient, "SD"  # restart along the gradient
            if not self._line_search():
                return False
        denom = np.dot(self.gradient_old, self.gradient_old)
        beta = max(0.0, np.dot(self.gradient, self.gradient - self.gradient_old) / denom) if denom > 0.0 else 0.0
        self.direction = -self.gradient + beta * self.direction
        self.status = "CG" if beta > 0.0 else "SD"
        return True


class CGOptimizer(BaseOptimizer):
    """Conjugate gradients with a Newton line search (opt.py:147-180).  The reference delegates the algorithm to
    ``molmod.minimizer``; when molmod is importable that implementation is used (identical iterates), otherwise the
    self-contained ``_CGMinimizer`` above - so ``simulations/relaxed_scan.py`` runs through the drop-in either way."""

    log_name = "CGOPT"

    def __init__(self, dof, state=None, hooks=None, counter0=0):
        try:
            from molmod.minimizer import ConjugateGradient, NewtonLineSearch, Minimizer

            if not hasattr(Minimizer, "propagate"):  # the test stand-in of molmod has no algorithm behind it
                raise ImportError("molmod.minimizer stub")
            self.minimizer = Minimizer(dof.x0, self.fun, ConjugateGradient(), NewtonLineSearch(), None, None,
                                       anagrad=True, verbose=False)
        except ImportError:
            self.minimizer = _CGMinimizer(dof.x0, self.fun)
        BaseOptimizer.__init__(self, dof, state, hooks, counter0)

    def initialize(self):
        self.minimizer.initialize()
        BaseOptimizer.initialize(self)

    def propagate(self):
        success = self.minimizer.propagate()
        self.x = self.minimizer.x
        if not success:
            return True  # line search failed: give up, like the reference
        return BaseOptimizer.propagate(self)


class HessianModel(object):
    """Dense model of the Hessian, identity unless ``hessian0`` is given; subclasses implement the secant ``update``
    and return False when they had to skip it."""

    def __init__(self, ndof, hessian0=None):
        self.ndof = ndof
        self.hessian = np.identity(ndof, float) if hessian0 is None else hessian0.copy()
        if self.hessian.shape != (ndof, ndof):
            raise TypeError("Incorrect shape of the initial hessian in quasi-newton method.")

    def get_spectrum(self):
        return np.linalg.eigh(self.hessian)

    def _add_rank_one(self, vec, scale):
        self.hessian += np.outer(vec, vec) / scale


class BFGSHessianModel(HessianModel):
    def update(self, dx, dg):
        """Rank-two update; skipped when either curvature is not safely positive (opt.py:208-236)."""
        hdx = np.dot(self.hessian, dx)
        hmax = abs(self.hessian).max()
        curv_model, curv_true = np.dot(dx, hdx), np.dot(dg, dx)
        if hmax * curv_model <= 1e-5 * abs(hdx).max() ** 2 or hmax * curv_true <= 1e-5 * abs(dg).max() ** 2:
            return False
        self._add_rank_one(hdx, -curv_model)
        self._add_rank_one(dg, curv_true)
        return True


class SR1HessianModel(HessianModel):
    def update(self, dx, dg):
        """Symmetric rank-one update; skipped when the denominator is tiny (opt.py:239-255)."""
        resid = dg - np.dot(self.hessian, dx)
        denom = np.dot(resid, dx)
        if abs(denom) <= 1e-5 * np.linalg.norm(dx) * np.linalg.norm(resid):
            return False
        self._add_rank_one(resid, denom)
        return True


class QNOptimizer(BaseOptimizer):
    """Trust-radius quasi-Newton optimiser (opt.py:258-393): SR1 model, step from the model's spectrum inside the trust
    radius, acceptance on an energy decrease (on a gradient-norm decrease once the radius is below ``small_radius``)."""

    log_name = "QNOPT"

    def __init__(self, dof, state=None, hooks=None, counter0=0, trust_radius=1.0, small_radius=1e-5,
                 too_small_radius=1e-10, hessian0=None):
        self.x_old = dof.x0
        self.hessian = SR1HessianModel(len(dof.x0), hessian0)
        self.trust_radius = self.initial_trust_radius = trust_radius
        self.small_radius, self.too_small_radius = small_radius, too_small_radius
        BaseOptimizer.__init__(self, dof, state, hooks, counter0)

    def _advance(self):
        self.x, self.f, self.g = self.make_step()

    def initialize(self):
        self.f_old, self.g_old = self.fun(self.dof.x0, True)
        self._advance()
        BaseOptimizer.initialize(self)

    def propagate(self):
        if not self.hessian.update(self.x - self.x_old, self.g - self.g_old):
            # a failed update poisons the model: start over from the identity and the initial radius
            self.hessian = SR1HessianModel(len(self.x))
            self.trust_radius = self.initial_trust_radius
        self.x_old, self.f_old, self.g_old = self.x, self.f, self.g
        self._advance()
        return BaseOptimizer.propagate(self)

    def _rejects(self, f, g):
        """Energy went up - or, at radii where energy differences drown in rounding, the gradient norm did."""
        if f - self.f_old > 0:
            return True
        return self.trust_radius < self.small_radius and np.linalg.norm(g) - np.linalg.norm(self.g_old) > 0

    def _shrink_below(self, radius):
        """Halve the trust radius until it is strictly inside the step that was just rejected (opt.py:376-382)."""
        self.trust_radius *= 0.5
        while self.trust_radius >= radius:
            self.trust_radius *= 0.5
        if self.trust_radius < self.too_small_radius:
            raise RuntimeError("The trust radius becomes too small. Is the potential energy surface smooth?")

    def make_step(self):
        evals, evecs = self.hessian.get_spectrum()
        grad_eigen = np.dot(evecs.T, self.g_old)
        while True:
            delta_eigen = solve_trust_radius(grad_eigen, evals, self.trust_radius)
            x = self.x_old + np.dot(evecs, delta_eigen)
            f, g = self.fun(x, True)
            if self._rejects(f, g):
                self._shrink_below(np.linalg.norm(delta_eigen))
                continue
            if self.trust_radius < self.initial_trust_radius:
                self.trust_radius *= 2.0
            return x, f, g


def solve_trust_radius(grad, evals, radius, threshold=1e-5):
    """Step in the eigenbasis of the model Hessian that minimises the quadratic model inside the sphere of the given
    radius (opt.py:396-443): the Newton step when it fits (positive definite model), otherwise -grad / (evals + ridge)
    with the ridge found by a secant iteration on |step| - radius."""
    if evals.min() > 0:
        step = -grad / evals
        if np.linalg.norm(step) <= radius:
            return step

    def excess(ridge):
        return np.linalg.norm(grad / (evals + ridge)) - radius

    ridge_min = -evals.min()

    def bracket(alpha, sign):
        # walk away from (sign > 0) or towards (sign < 0) the pole at ridge_min until the excess has the wanted sign
        while True:
            ridge = ridge_min + alpha
            err = excess(ridge)
            if sign * err < 0:
                return ridge, err
            alpha = alpha * 2 if sign > 0 else alpha / 2

    a = bracket(min(1e1, abs(evals.max())), -1)
    b = bracket(max(1e-5, abs(ridge_min)), 1)
    err = np.inf
    while abs(err) > radius * threshold:
        ridge = (a[1] * a[0] - b[1] * b[0]) / (a[1] - b[1])
        err = excess(ridge)
        if err > 0 and a[1] > 0:
            a = (ridge, err)
        else:
            b = (ridge, err)
    return -grad / (evals + ridge)
