"""Host-side mirror of the parts of NamedTrajectories.NamedTrajectory that DirectTrajOpt.jl's hot
path reads: ``data`` is ``dim x N``, ``datavec = vec(data)`` is knot-major,
``components[name]`` is the index range of a component inside one knot (0-based here, 1-based in
Julia), ``traj[k]`` is the 1-based knot point."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


class KnotPoint:
    def __init__(self, traj, k):
        self.traj, self.k = traj, k
        self.data = traj.data[:, k - 1]
        self.components = traj.components

    def __getitem__(self, name):
        return self.data[self.traj.components[name]]

    @property
    def timestep(self):
        return float(self.data[self.traj.components[self.traj.timestep]][0])


class NamedTrajectory:
    """``NamedTrajectory(components; controls, timestep, bounds, initial, final, goal)``.

    ``components`` maps name -> array ``(dim, N)`` (1-D arrays are treated as ``(1, N)``), in the
    order the knot is laid out.  ``timestep`` names the free time-step component (the reference's
    ``timestep::Symbol``)."""

    def __init__(self, components, controls=(), timestep="dt", bounds=None, initial=None, final=None, goal=None,
                 global_components=None):
        comps = OrderedDict()
        for name, arr in components.items():
            a = np.asarray(arr, dtype=np.float64)
            if a.ndim == 1:
                a = a[None, :]
            comps[name] = a
        Ns = {a.shape[1] for a in comps.values()}
        if len(Ns) != 1:
            raise ValueError("all components must have the same number of knots")
        self.N = Ns.pop()
        if not isinstance(timestep, str) or timestep not in comps:
            raise ValueError("timestep must name a component (variable time step); fixed time steps are not in scope")
        self.timestep = timestep
        self.names = tuple(comps)
        self.dims = {n: a.shape[0] for n, a in comps.items()}
        self.components = {}
        off = 0
        for n, a in comps.items():
            self.components[n] = range(off, off + a.shape[0])
            off += a.shape[0]
        self.dim = off
        self.data = np.asfortranarray(np.vstack([comps[n] for n in self.names]))
        self.control_names = tuple(controls)
        self.bounds = dict(bounds or {})
        self.initial = dict(initial or {})
        self.final = dict(final or {})
        self.goal = {k: np.asarray(v, float) for k, v in (goal or {}).items()}
        # global (non-time-varying) variables: appended to the knots in the solver's vector
        # (NamedTrajectories' global_data / global_components; evaluator.jl:474-482)
        self.global_components, self.global_dims = {}, {}
        gvals, off = [], 0
        for name, arr in (global_components or {}).items():
            a = np.atleast_1d(np.asarray(arr, dtype=np.float64)).reshape(-1)
            self.global_components[name] = range(off, off + a.size)
            self.global_dims[name] = a.size
            gvals.append(a)
            off += a.size
        self.global_names = tuple(self.global_components)
        self.global_dim = off
        self.global_data = np.concatenate(gvals) if gvals else np.zeros(0)

    @property
    def datavec(self):
        return self.data.reshape(-1, order="F")

    def vec(self):
        """``vec(traj)`` = [datavec; global_data], the solver's primal vector."""
        return np.concatenate([self.datavec, self.global_data])

    def __getitem__(self, k):
        if not 1 <= k <= self.N:
            raise IndexError(k)
        return KnotPoint(self, k)

    def __getattr__(self, name):
        comps = self.__dict__.get("components", {})
        if name in comps:
            return self.data[comps[name], :]
        raise AttributeError(name)

    def copy_with(self, datavec):
        t = object.__new__(NamedTrajectory)
        t.__dict__.update(self.__dict__)
        v = np.asarray(datavec, float)
        t.data = np.asfortranarray(v[: self.dim * self.N].reshape(self.dim, self.N, order="F").copy())
        if v.size > self.dim * self.N:
            t.global_data = v[self.dim * self.N :].copy()
        return t

    def update(self, Z):
        """NamedTrajectories.update!(traj, Z; type=:both)."""
        Z = np.asarray(Z, float)
        self.data[...] = Z[: self.dim * self.N].reshape(self.dim, self.N, order="F")
        if self.global_dim:
            self.global_data[...] = Z[self.dim * self.N : self.dim * self.N + self.global_dim]
