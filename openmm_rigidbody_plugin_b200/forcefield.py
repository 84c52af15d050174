"""Rigid-body extension of OpenMM's app.ForceField: the part of the reference's Python layer that PRODUCES the
`bodyIndices` the integrator consumes (python/forcefield.py:3-178 of the reference: registerBodyTemplate,
resolveBodies with mergeList, removeConstraints, removeForces, createSystem -> (system, bodyIndices)).

When OpenMM's application layer is importable, `ForceField` subclasses it exactly like the reference does.
It is not installed in this environment, so the body-resolution logic lives in a mixin that only relies on
duck-typed objects (a topology with residues()/atoms()/getNumAtoms(), residue templates with atoms and
virtualSites, a system with constraints and forces) and is tested with fake topologies.
"""
from __future__ import annotations

import re


class _BodyTemplate:
    """A registered rigid-body template: atoms of residue `residue` whose names match `pattern`."""

    def __init__(self, residue, pattern, virtualSites):
        self.residue = residue
        self.pattern = pattern
        self.virtualSites = virtualSites

    def __str__(self):
        return "%s: [%s]" % (self.residue, self.pattern)


def _is_sequence(x):
    return hasattr(x, "__iter__") and not isinstance(x, (str, bytes))


class RigidBodyForceFieldMixin:
    """Body-template bookkeeping; expects the host class to provide `_templates` (residue name -> template with
    `.atoms[i].name` and `.virtualSites[k].index`) and `createSystem(topology, **kwargs)`."""

    def _bodies(self):
        if not hasattr(self, "_bodyTemplates"):
            self._bodyTemplates = {}
        return self._bodyTemplates

    def getBodyTemplate(self, name):
        return self._bodies()[name]

    def registerBodyTemplate(self, name, residue, pattern=".+"):
        """Register a rigid-body template `name` made of the atoms of residue template `residue` whose names
        match the regular expression `pattern` (virtual sites never belong to a body)."""
        if name in self._bodies():
            raise ValueError("A rigid body template named %s has already been registered." % name)
        if residue not in self._templates:
            raise ValueError("Unknown residue %s in rigid body registration." % residue)
        template = self._templates[residue]
        names = [a.name for a in template.atoms]
        virtual = set(names[vs.index] for vs in template.virtualSites)
        self._bodies()[name] = _BodyTemplate(residue, pattern, virtual)

    @staticmethod
    def _disjointSets(sets):
        """Union overlapping sets until all are pairwise disjoint (transitive closure of 'shares a member')."""
        groups = []
        for s in sets:
            s = set(s)
            touching = [g for g in groups if not g.isdisjoint(s)]
            for g in touching:
                s |= g
                groups.remove(g)
            groups.append(s)
        return groups

    def resolveBodies(self, topology, merge=None):
        """Body index of every atom of `topology` (0 = free atom).  Bodies are numbered 1, 2, ... per (residue,
        matching template) in topology order; `merge` is a list of body indices - or a list of such lists - to be
        fused, the fused body taking the smallest index of its group (this leaves gaps, which the integrator's
        index cleaning compacts)."""
        index = [0] * topology.getNumAtoms()
        templates = list(self._bodies().values())
        n = 0
        for res in topology.residues():
            for body in templates:
                if body.residue != res.name:
                    continue
                n += 1
                atoms = [a for a in res.atoms() if a.name not in body.virtualSites and re.match(body.pattern, a.name)]
                if not atoms:
                    raise ValueError("no atom in residue %s matches pattern %s" % (res.name, body.pattern))
                for a in atoms:
                    index[a.index] = n
        if merge is not None:
            if not _is_sequence(merge):
                raise ValueError("merge parameter is not a sequence")
            merge = list(merge)
            groups = self._disjointSets(merge) if all(_is_sequence(m) for m in merge) else [set(merge)]
            target = {}
            for g in groups:
                lowest = min(g)
                for b in g:
                    target[b] = lowest
            index = [target.get(b, b) for b in index]
        return index

    @staticmethod
    def _intraBodyPairs(bodyIndices):
        """All atom pairs (i, j), i > j, that belong to the same rigid body."""
        members = {}
        for i, b in enumerate(bodyIndices):
            if b != 0:
                members.setdefault(b, []).append(i)
        pairs = []
        for b in sorted(members):
            atoms = members[b][::-1]
            for k, i in enumerate(atoms):
                for j in atoms[k + 1:]:
                    pairs.append((i, j))
        return pairs

    @staticmethod
    def removeConstraints(system, bodyIndices):
        """Silently drop every constraint that involves a rigid-body atom (the integrator rejects them)."""
        for i in reversed(range(system.getNumConstraints())):
            atom1, atom2 = system.getConstraintParameters(i)[:2]
            if bodyIndices[atom1] != 0 or bodyIndices[atom2] != 0:
                system.removeConstraint(i)

    def removeForces(self, system, bodyIndices):
        """Zero the non-bonded interactions between atoms of the same body (addException(i, j, 0, 1, 0, replace=True)
        on every force that looks like a NonbondedForce)."""
        forces = [system.getForce(i) for i in range(system.getNumForces())]
        nonbonded = [f for f in forces if self._isNonbonded(f)]
        for (i, j) in self._intraBodyPairs(bodyIndices):
            for force in nonbonded:
                force.addException(i, j, 0, 1, 0, replace=True)

    @staticmethod
    def _isNonbonded(force):
        return type(force).__name__ in ("NonbondedForce", "CustomNonbondedForce") or getattr(force, "isNonbonded", False)

    def createSystem(self, topology, **kwargs):
        """`ForceField.createSystem` plus rigid bodies: returns (system, bodyIndices).  Extra keyword arguments:
        mergeList (nested list of body indices to fuse), removeConstraints, removeForces (bool)."""
        mergeList = kwargs.pop("mergeList", None)
        removeForces = kwargs.pop("removeForces", False)
        removeConstraints = kwargs.pop("removeConstraints", False)
        system = super().createSystem(topology, **kwargs)
        bodyIndices = self.resolveBodies(topology, merge=mergeList)
        if removeConstraints:
            self.removeConstraints(system, bodyIndices)
        if removeForces:
            self.removeForces(system, bodyIndices)
        return (system, bodyIndices)


try:  # the real thing, when OpenMM is installed
    try:
        from openmm import app as _app
    except ImportError:
        from simtk.openmm import app as _app

    class ForceField(RigidBodyForceFieldMixin, _app.ForceField):
        def __init__(self, *files):
            super().__init__(*files)
            self._bodyTemplates = {}

except ImportError:
    _app = None

    class ForceField(RigidBodyForceFieldMixin):
        """Stand-alone variant (no OpenMM): residue templates are supplied directly and `createSystem` delegates to a
        user-provided builder `system_builder(topology, **kwargs)`."""

        def __init__(self, templates=None, system_builder=None):
            self._templates = dict(templates or {})
            self._bodyTemplates = {}
            self._builder = system_builder

        def _buildSystem(self, topology, **kwargs):
            if self._builder is None:
                raise RuntimeError("OpenMM is not installed: pass system_builder= to ForceField to create systems")
            return self._builder(topology, **kwargs)

        def createSystem(self, topology, **kwargs):
            mergeList = kwargs.pop("mergeList", None)
            removeForces = kwargs.pop("removeForces", False)
            removeConstraints = kwargs.pop("removeConstraints", False)
            system = self._buildSystem(topology, **kwargs)
            bodyIndices = self.resolveBodies(topology, merge=mergeList)
            if removeConstraints:
                self.removeConstraints(system, bodyIndices)
            if removeForces:
                self.removeForces(system, bodyIndices)
            return (system, bodyIndices)
