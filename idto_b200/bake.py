"""Bake step: URDF -> flat model tables (the BakedModel the CUDA layer consumes).

The reference queries Drake's MultibodyPlant/SceneGraph on every inverse
dynamics evaluation (optimizer/trajectory_optimizer.cc:228-386).  Drake is not
available here, and on a GPU the right design is to query the
plant ONCE and bake tree / inertia / contact-pair tables into device buffers.
This module is that one-time query, restated for URDF input in Drake's
conventions:

  * moving bodies in depth-first dof order, trees in joint order;
  * a free root link gets a quaternion floating joint  q=[qw qx qy qz x y z],
    v=[w_W, v_W];
  * `fixed` joints weld: the child is merged into its parent (inertia + geometry);
  * revolute/continuous/prismatic: F = joint origin on the parent, M = B;
  * planar: Fz = URDF axis (plane normal), q = [x_F, y_F, theta];
  * collision geometry ids in registration (file) order; candidate pairs after
    Drake's default filtering (same body, adjacent bodies, anchored-anchored,
    explicit drake:collision_filter_group), sorted by (idA, idB) with A < B;
  * actuated dofs from <transmission> (skipped when the joint's effort limit is 0).

Nothing here runs on the hot path.
"""
from __future__ import annotations

import ctypes
import json
import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_PLANAR, JOINT_QUAT_FLOATING = 0, 1, 2, 3
GEOM_SPHERE, GEOM_BOX, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_HALF_SPACE = 0, 1, 2, 3, 4
_GEOM_CODE = {"sphere": GEOM_SPHERE, "box": GEOM_BOX, "capsule": GEOM_CAPSULE, "cylinder": GEOM_CYLINDER,
              "half_space": GEOM_HALF_SPACE}
_JOINT_NQ = {JOINT_REVOLUTE: 1, JOINT_PRISMATIC: 1, JOINT_PLANAR: 3, JOINT_QUAT_FLOATING: 7}
_JOINT_NV = {JOINT_REVOLUTE: 1, JOINT_PRISMATIC: 1, JOINT_PLANAR: 3, JOINT_QUAT_FLOATING: 6}


# ----------------------------------------------------------------------------
# small rigid-transform helpers (host side, numpy)
def rpy_to_R(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


class X:
    """Rigid transform (R, p)."""

    def __init__(self, R=None, p=None):
        self.R = np.eye(3) if R is None else np.asarray(R, float)
        self.p = np.zeros(3) if p is None else np.asarray(p, float)

    def __matmul__(self, o):
        if isinstance(o, X):
            return X(self.R @ o.R, self.p + self.R @ o.p)
        return self.R @ np.asarray(o, float) + self.p

    def inv(self):
        return X(self.R.T, -self.R.T @ self.p)

    def flat(self):
        return np.concatenate([self.R.reshape(-1), self.p])


def make_from_one_vector(u, axis_index):
    """Right-handed basis with column `axis_index` = u/|u|.

    Restates Drake's RotationMatrix::MakeFromOneVector construction
    (v = a x u/|a x u|, w = u x v with `a` the unit vector of u's smallest
    component).  Ties between equally small components are resolved toward the
    LAST index: that is the choice consistent with the reference's hopper
    example, whose q = [height, horizontal, theta] (examples/hopper/hopper.yaml:7-8
    with models/hopper.urdf:36-41) — parity unpinned otherwise (no Drake here).
    """
    u = np.asarray(u, float)
    u = u / np.linalg.norm(u)
    au = np.abs(u)
    i = int(max(k for k in range(3) if au[k] == au.min()))
    j, k = (i + 1) % 3, (i + 2) % 3
    r = math.hypot(u[j], u[k])
    s = 1.0 / r
    v = np.zeros(3)
    w = np.zeros(3)
    v[j], v[k] = -u[k] * s, u[j] * s
    w[i], w[j], w[k] = r, -u[i] * u[j] * s, -u[i] * u[k] * s
    R = np.zeros((3, 3))
    R[:, axis_index] = u
    R[:, (axis_index + 1) % 3] = v
    R[:, (axis_index + 2) % 3] = w
    return R


# ----------------------------------------------------------------------------
@dataclass
class _Link:
    name: str
    mass: float = 0.0
    com: np.ndarray = field(default_factory=lambda: np.zeros(3))
    I_cm: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))  # about com, link axes
    collisions: list = field(default_factory=list)  # (reg_id, type, dims, X_LG)


@dataclass
class _Joint:
    name: str
    type: str
    parent: str
    child: str
    X_PJ: X
    axis: np.ndarray
    damping: np.ndarray
    effort: float | None
    index: int


@dataclass
class BakedModel:
    """Flat tables; field names follow include/idto_b200.h:idto_model_desc."""
    name: str
    nbodies: int
    nq: int
    nv: int
    parent: np.ndarray
    joint_type: np.ndarray
    q_start: np.ndarray
    v_start: np.ndarray
    X_PF: np.ndarray
    R_MB: np.ndarray
    axis: np.ndarray
    damping: np.ndarray
    mass: np.ndarray
    com: np.ndarray
    inertia: np.ndarray
    gravity: np.ndarray
    actuated: np.ndarray
    ngeoms: int
    geom_body: np.ndarray
    geom_type: np.ndarray
    geom_dims: np.ndarray
    X_BG: np.ndarray
    npairs: int
    pair_geomA: np.ndarray
    pair_geomB: np.ndarray
    body_names: list = field(default_factory=list)
    geom_names: list = field(default_factory=list)
    quat_q_starts: list = field(default_factory=list)

    _INT = ("parent", "joint_type", "q_start", "v_start", "actuated", "geom_body", "geom_type",
            "pair_geomA", "pair_geomB")
    _DBL = ("X_PF", "R_MB", "axis", "damping", "mass", "com", "inertia", "gravity", "geom_dims", "X_BG")

    @property
    def unactuated_dofs(self):
        """Reference: trajectory_optimizer.cc:63-72 (all actuated if B is empty)."""
        if not np.any(self.actuated):
            return []
        return [i for i in range(self.nv) if not self.actuated[i]]

    # -- (de)serialisation: the baked tables are what travels to the GPU box
    def to_json(self):
        d = {"name": self.name, "nbodies": self.nbodies, "nq": self.nq, "nv": self.nv,
             "ngeoms": self.ngeoms, "npairs": self.npairs, "body_names": self.body_names,
             "geom_names": self.geom_names, "quat_q_starts": self.quat_q_starts}
        for k in self._INT:
            d[k] = np.asarray(getattr(self, k)).astype(int).reshape(-1).tolist()
        for k in self._DBL:
            d[k] = [float.hex(float(x)) for x in np.asarray(getattr(self, k), float).reshape(-1)]
        return d

    @classmethod
    def from_json(cls, d):
        kw = {k: d[k] for k in ("name", "nbodies", "nq", "nv", "ngeoms", "npairs", "body_names",
                                "geom_names", "quat_q_starts")}
        for k in cls._INT:
            kw[k] = np.asarray(d[k], dtype=np.int32)
        for k in cls._DBL:
            kw[k] = np.asarray([float.fromhex(x) for x in d[k]], dtype=np.float64)
        nb, ng = kw["nbodies"], kw["ngeoms"]
        kw["X_PF"] = kw["X_PF"].reshape(nb, 12)
        kw["R_MB"] = kw["R_MB"].reshape(nb, 9)
        kw["axis"] = kw["axis"].reshape(nb, 3)
        kw["com"] = kw["com"].reshape(nb, 3)
        kw["inertia"] = kw["inertia"].reshape(nb, 6)
        kw["geom_dims"] = kw["geom_dims"].reshape(ng, 3)
        kw["X_BG"] = kw["X_BG"].reshape(ng, 12)
        return cls(**kw)

    def save(self, path):
        with open(path, "w") as f:
            json.dump(self.to_json(), f)

    def save_txt(self, path):
        """Whitespace-separated table dump read by include/idto_b200.hpp (MultibodyPlant::LoadBaked)."""
        with open(path, "w") as f:
            f.write(f"{self.nbodies} {self.nq} {self.nv} {self.ngeoms} {self.npairs}\n")
            for k in ("parent", "joint_type", "q_start", "v_start", "actuated", "geom_body", "geom_type",
                      "pair_geomA", "pair_geomB"):
                f.write(" ".join(str(int(x)) for x in np.asarray(getattr(self, k)).reshape(-1)) + "\n")
            for k in ("X_PF", "R_MB", "axis", "damping", "mass", "com", "inertia", "gravity", "geom_dims", "X_BG"):
                f.write(" ".join(repr(float(x)) for x in np.asarray(getattr(self, k), float).reshape(-1)) + "\n")

    @classmethod
    def load(cls, path):
        with open(path) as f:
            return cls.from_json(json.load(f))


class ModelDesc(ctypes.Structure):
    """ctypes mirror of idto_model_desc (include/idto_b200.h)."""
    _I, _D = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    _fields_ = [("nbodies", ctypes.c_int), ("nq", ctypes.c_int), ("nv", ctypes.c_int),
                ("parent", _I), ("joint_type", _I), ("q_start", _I), ("v_start", _I),
                ("X_PF", _D), ("R_MB", _D), ("axis", _D), ("damping", _D), ("mass", _D),
                ("com", _D), ("inertia", _D), ("gravity", ctypes.c_double * 3), ("actuated", _I),
                ("ngeoms", ctypes.c_int), ("geom_body", _I), ("geom_type", _I), ("geom_dims", _D),
                ("X_BG", _D), ("npairs", ctypes.c_int), ("pair_geomA", _I), ("pair_geomB", _I)]


def model_desc(m: BakedModel):
    """Returns (ModelDesc, keepalive) for passing a BakedModel across the C ABI."""
    keep = {}
    d = ModelDesc()
    d.nbodies, d.nq, d.nv, d.ngeoms, d.npairs = m.nbodies, m.nq, m.nv, m.ngeoms, m.npairs
    for k in BakedModel._INT:
        a = np.ascontiguousarray(getattr(m, k), dtype=np.int32).reshape(-1)
        if a.size == 0:
            a = np.zeros(1, np.int32)
        keep[k] = a
        setattr(d, k, a.ctypes.data_as(ModelDesc._I))
    for k in BakedModel._DBL:
        a = np.ascontiguousarray(getattr(m, k), dtype=np.float64).reshape(-1)
        if a.size == 0:
            a = np.zeros(1, np.float64)
        keep[k] = a
        if k == "gravity":
            d.gravity = (ctypes.c_double * 3)(*a.tolist())
        else:
            setattr(d, k, a.ctypes.data_as(ModelDesc._D))
    return d, keep


# ----------------------------------------------------------------------------
def _vec(s, n=3, default=None):
    if s is None:
        return np.array(default if default is not None else [0.0] * n, float)
    return np.array([float(t) for t in s.split()], float)


def _origin(node):
    o = node.find("origin") if node is not None else None
    if o is None:
        return X()
    xyz = _vec(o.get("xyz"))
    rpy = _vec(o.get("rpy"))
    return X(rpy_to_R(*rpy), xyz)


class ModelBuilder:
    """Minimal stand-in for `MultibodyPlant` + `Parser(plant).AddModels(urdf)`.

    Usage mirrors the reference examples (e.g. examples/hopper/hopper.cc:40-50):
        b = ModelBuilder(); b.add_urdf(path)
        b.register_collision_geometry("world", X(p=[0,0,-5]), "box", [25,25,10], "ground")
        model = b.finalize()
    """

    def __init__(self, name="model"):
        self.name = name
        self.links: dict[str, _Link] = {"world": _Link("world")}
        self.link_order = ["world"]
        self.joints: list[_Joint] = []
        self.actuated_joints: set[str] = set()
        self.filter_groups: dict[str, set] = {}
        self.filter_excludes: list[tuple] = []
        self._next_geom = 0
        self.gravity = np.array([0.0, 0.0, -9.81])

    # -- parsing --------------------------------------------------------------
    def add_urdf(self, path):
        with open(path) as f:
            text = f.read()
        if "drake:" in text and "xmlns:drake" not in text:
            # Drake's URDF parser tolerates the unbound `drake:` prefix; expat does not.
            text = text.replace("<robot ", '<robot xmlns:drake="http://drake.mit.edu" ', 1)
        root = ET.fromstring(text)
        self.name = root.get("name", self.name)
        for ln in root.findall("link"):
            link = _Link(ln.get("name"))
            inr = ln.find("inertial")
            if inr is not None:
                Xi = _origin(inr)
                m = inr.find("mass")
                link.mass = float(m.get("value")) if m is not None else 0.0
                link.com = Xi.p.copy()
                it = inr.find("inertia")
                if it is not None:
                    g = lambda k: float(it.get(k, "0"))
                    I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")],
                                  [g("ixz"), g("iyz"), g("izz")]])
                    link.I_cm = Xi.R @ I @ Xi.R.T
            for col in ln.findall("collision"):
                geo = col.find("geometry")
                Xg = _origin(col)
                shape = None
                for child in geo:
                    tag = child.tag.split("}")[-1]
                    if tag == "sphere":
                        shape = ("sphere", [float(child.get("radius")), 0.0, 0.0])
                    elif tag == "box":
                        shape = ("box", list(_vec(child.get("size"))))
                    elif tag in ("capsule", "cylinder"):  # axis = z of the geometry frame, centred (Drake / URDF)
                        shape = (tag, [float(child.get("radius")), float(child.get("length")), 0.0])
                    else:
                        shape = (tag, [0.0, 0.0, 0.0])
                    break
                link.collisions.append((self._next_geom, shape[0], shape[1], Xg,
                                        col.get("name", f"{link.name}_collision{len(link.collisions)}")))
                self._next_geom += 1
            self.links[link.name] = link
            self.link_order.append(link.name)
        for jn in root.findall("joint"):
            ax = jn.find("axis")
            axis = _vec(ax.get("xyz")) if ax is not None else np.array([1.0, 0.0, 0.0])
            if np.linalg.norm(axis) > 0:
                axis = axis / np.linalg.norm(axis)
            dyn = jn.find("dynamics")
            damping = _vec(dyn.get("damping")) if dyn is not None and dyn.get("damping") else np.zeros(1)
            lim = jn.find("limit")
            effort = float(lim.get("effort")) if lim is not None and lim.get("effort") is not None else None
            self.joints.append(_Joint(jn.get("name"), jn.get("type"), jn.find("parent").get("link"),
                                      jn.find("child").get("link"), _origin(jn), axis, damping, effort,
                                      len(self.joints)))
        for tr in root.findall("transmission"):
            j = tr.find("joint")
            if j is not None:
                self.actuated_joints.add(j.get("name"))
        for el in root:
            if el.tag.split("}")[-1].endswith("collision_filter_group") and "ignored" not in el.tag:
                gname = el.get("name")
                members, ignored = set(), []
                for ch in el:
                    t = ch.tag.split("}")[-1]
                    if t.endswith("member"):
                        members.add(ch.get("link"))
                    elif t.endswith("ignored_collision_filter_group"):
                        ignored.append(ch.get("name"))
                self.filter_groups[gname] = members
                for ig in ignored:
                    self.filter_excludes.append((gname, ig))
        return self

    def add_sdf(self, path):
        """`Parser(plant).AddModels(sdf)` analogue for single-model SDF 1.7 files (models/allegro_hand.sdf).

        SDF semantics restated: a link `<pose>` is expressed in the model frame; a joint frame coincides with
        its child link frame (no joint `<pose>` in this file); `<axis><xyz expressed_in="__model__">` is given
        in the model frame; `<inertial><pose>` places the centre of mass and the inertia axes in the link
        frame; joints with a non-zero effort limit get an actuator (Drake's SDF parser).  When a collision
        `<geometry>` lists several shapes (palm: `<sphere>` and `<box>`, allegro_hand.sdf:47-56) sdformat's
        Geometry::Load keeps the first match of its fixed order box, capsule, cylinder, ellipsoid, plane,
        sphere — the box."""
        with open(path) as f:
            text = f.read()
        if "drake:" in text and "xmlns:drake" not in text:
            text = text.replace("<sdf ", '<sdf xmlns:drake="http://drake.mit.edu" ', 1)
        model = ET.fromstring(text).find("model")
        self.name = model.get("name", self.name)

        def pose(node):
            pn = node.find("pose") if node is not None else None
            if pn is None or not (pn.text or "").strip():
                return X()
            v = [float(t) for t in pn.text.split()]
            return X(rpy_to_R(*v[3:6]), v[:3])

        X_ML = {}
        for ln in model.findall("link"):
            link = _Link(ln.get("name"))
            X_ML[link.name] = pose(ln)
            inr = ln.find("inertial")
            if inr is not None:
                Xi = pose(inr)
                link.mass = float(inr.findtext("mass", "0"))
                link.com = Xi.p.copy()
                it = inr.find("inertia")
                if it is not None:
                    g = lambda k: float(it.findtext(k, "0"))
                    I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")],
                                  [g("ixz"), g("iyz"), g("izz")]])
                    link.I_cm = Xi.R @ I @ Xi.R.T
            for col in ln.findall("collision"):
                geo = col.find("geometry")
                shape = None
                for tag in ("box", "capsule", "cylinder", "ellipsoid", "plane", "sphere"):
                    ch = geo.find(tag)
                    if ch is None:
                        continue
                    if tag == "box":
                        shape = ("box", [float(t) for t in ch.findtext("size").split()])
                    elif tag == "sphere":
                        shape = ("sphere", [float(ch.findtext("radius")), 0.0, 0.0])
                    elif tag in ("capsule", "cylinder"):
                        shape = (tag, [float(ch.findtext("radius")), float(ch.findtext("length")), 0.0])
                    else:
                        shape = (tag, [0.0, 0.0, 0.0])
                    break
                link.collisions.append((self._next_geom, shape[0], shape[1], pose(col),
                                        col.get("name", f"{link.name}_collision{len(link.collisions)}")))
                self._next_geom += 1
            self.links[link.name] = link
            self.link_order.append(link.name)
        for jn in model.findall("joint"):
            parent, child = jn.findtext("parent"), jn.findtext("child")
            ax = jn.find("axis")
            axis_in, frame = np.array([0.0, 0.0, 1.0]), ""
            damping, effort = np.zeros(1), None
            if ax is not None:
                xyz = ax.find("xyz")
                axis_in = np.array([float(t) for t in xyz.text.split()])
                frame = xyz.get("expressed_in", "")
                dyn = ax.find("dynamics")
                if dyn is not None and dyn.findtext("damping"):
                    damping = np.array([float(dyn.findtext("damping"))])
                lim = ax.find("limit")
                if lim is not None and lim.findtext("effort"):
                    effort = float(lim.findtext("effort"))
            X_MC, X_MP = X_ML[child], X_ML.get(parent, X())
            axis = X_MC.R.T @ axis_in if frame == "__model__" else axis_in  # joint frame == child link frame
            axis = axis / np.linalg.norm(axis)
            jtype = {"revolute": "revolute", "prismatic": "prismatic", "fixed": "fixed"}[jn.get("type")]
            self.joints.append(_Joint(jn.get("name"), jtype, parent, child, X_MP.inv() @ X_MC, axis, damping, effort,
                                      len(self.joints)))
            if jtype != "fixed" and not (effort is not None and effort == 0.0):
                self.actuated_joints.add(jn.get("name"))
        self._sdf_root_links = [n for n in X_ML if n not in {j.child for j in self.joints}]
        self._sdf_X_ML = X_ML
        return self

    def weld_frames(self, parent_link, child_link, X_PC):
        """plant.WeldFrames(parent_frame, child_frame, X_PC) analogue (a fixed joint)."""
        self.joints.append(_Joint(f"weld_{child_link}", "fixed", parent_link, child_link, X_PC, np.array([0.0, 0.0, 1.0]),
                                  np.zeros(1), None, len(self.joints)))
        return self

    def add_rigid_body(self, name, mass, com, I_cm):
        """plant.AddRigidBody(name, SpatialInertia) analogue: a free body gets a quaternion floating joint at
        Finalize (after every declared joint)."""
        link = _Link(name)
        link.mass, link.com, link.I_cm = float(mass), np.asarray(com, float), np.asarray(I_cm, float)
        self.links[name] = link
        self.link_order.append(name)
        return self

    def register_collision_geometry(self, link, X_LG, shape, dims, name="geom"):
        """plant.RegisterCollisionGeometry(body, X_BG, shape, name, ...) analogue."""
        d = list(dims) + [0.0] * (3 - len(dims))
        self.links[link].collisions.append((self._next_geom, shape, d, X_LG, name))
        self._next_geom += 1
        return self

    @staticmethod
    def half_space_pose(normal, point):
        """HalfSpace::MakePose(Hz_dir_F, p_FB) analogue: a frame whose z axis is `normal` and whose origin is
        `point` (any completion of the basis describes the same half space)."""
        z = np.asarray(normal, float) / np.linalg.norm(normal)
        a = np.eye(3)[int(np.argmin(np.abs(z)))]
        x = np.cross(a, z)
        x /= np.linalg.norm(x)
        return X(np.column_stack([x, np.cross(z, x), z]), np.asarray(point, float))

    # -- finalize -------------------------------------------------------------
    def finalize(self) -> BakedModel:
        child_joint = {j.child: j for j in self.joints}
        children: dict[str, list] = {n: [] for n in self.links}
        roots = []
        for n in self.link_order:
            if n == "world":
                continue
            if n in child_joint:
                children[child_joint[n].parent].append(child_joint[n])
            else:
                roots.append(n)  # free body -> quaternion floating joint (added at Finalize, last joint index)
        for n in children:
            children[n].sort(key=lambda j: j.index)

        bodies = []  # dict per moving body
        link_to_body = {"world": (-1, X())}  # link -> (moving body idx, X_B_link)

        def add_moving(link_name, parent_body, jtype, X_PF, R_MB, axis, damping):
            idx = len(bodies)
            bodies.append(dict(name=link_name, parent=parent_body, jtype=jtype, X_PF=X_PF, R_MB=R_MB,
                               axis=axis, damping=damping, mass=0.0, mc=np.zeros(3), I_o=np.zeros((3, 3)),
                               geoms=[]))
            link_to_body[link_name] = (idx, X())
            return idx

        def absorb(body_idx, link, X_BL):
            """Merge a link's inertia into moving body `body_idx` (frame offset X_BL)."""
            if body_idx < 0 or link.mass == 0.0 and not np.any(link.I_cm):
                return
            b = bodies[body_idx]
            c = X_BL @ link.com
            Icm = X_BL.R @ link.I_cm @ X_BL.R.T
            b["mass"] += link.mass
            b["mc"] += link.mass * c
            b["I_o"] += Icm + link.mass * (np.dot(c, c) * np.eye(3) - np.outer(c, c))

        geoms = []  # (reg_id, body_idx, type, dims, X_BG, name, link)

        def visit(link_name):
            """Depth-first over outboard joints of `link_name` in joint-index order."""
            body_idx, X_BL = link_to_body[link_name]
            link = self.links[link_name]
            absorb(body_idx, link, X_BL)
            for (gid, shape, dims, X_LG, gname) in link.collisions:
                geoms.append((gid, body_idx, shape, dims, X_BL @ X_LG, gname, link_name))
            for j in children[link_name]:
                X_BJ = X_BL @ j.X_PJ  # joint frame in the (merged) parent moving body
                if j.type == "fixed":
                    link_to_body[j.child] = (body_idx, X_BJ)
                    visit(j.child)
                    continue
                if j.type in ("revolute", "continuous"):
                    add_moving(j.child, body_idx, JOINT_REVOLUTE, X_BJ, np.eye(3), j.axis,
                               [float(j.damping[0])])
                elif j.type == "prismatic":
                    add_moving(j.child, body_idx, JOINT_PRISMATIC, X_BJ, np.eye(3), j.axis,
                               [float(j.damping[0])])
                elif j.type == "planar":
                    R_JF = make_from_one_vector(j.axis, 2)
                    d = list(j.damping) + [0.0] * (3 - len(j.damping)) if len(j.damping) == 3 else [0.0] * 3
                    # F = J*R on the parent, M = R on the child (so B == J at q = 0): R_MB = R_JF^T.
                    add_moving(j.child, body_idx, JOINT_PLANAR, X(X_BJ.R @ R_JF, X_BJ.p), R_JF.T,
                               np.array([0.0, 0.0, 1.0]), d)
                else:
                    raise NotImplementedError(f"joint type {j.type!r} ({j.name}) is outside the baked set")
                visit(j.child)

        # trees hanging off the world (incl. links welded to it), in joint order, then free bodies
        visit("world")
        for r in roots:
            add_moving(r, -1, JOINT_QUAT_FLOATING, X(), np.eye(3), np.array([0.0, 0.0, 1.0]), [0.0] * 6)
            visit(r)

        nb = len(bodies)
        q_start, v_start, nq, nv = [], [], 0, 0
        for b in bodies:
            q_start.append(nq)
            v_start.append(nv)
            nq += _JOINT_NQ[b["jtype"]]
            nv += _JOINT_NV[b["jtype"]]
        damping = np.zeros(nv)
        actuated = np.zeros(nv, np.int32)
        jbyname = {j.child: j for j in self.joints}
        for k, b in enumerate(bodies):
            n = _JOINT_NV[b["jtype"]]
            damping[v_start[k]:v_start[k] + n] = b["damping"][:n]
            j = jbyname.get(b["name"])
            if j is not None and j.name in self.actuated_joints and not (j.effort is not None and j.effort == 0.0):
                actuated[v_start[k]:v_start[k] + n] = 1
        # NB: Drake skips a transmission whose joint has a zero effort limit (acrobot shoulder,
        # models/acrobot/acrobot.urdf:38); the acrobot shoulder has no transmission anyway.

        mass = np.array([b["mass"] for b in bodies])
        com = np.array([b["mc"] / b["mass"] if b["mass"] > 0 else np.zeros(3) for b in bodies]).reshape(nb, 3)
        inertia = np.array([[b["I_o"][0, 0], b["I_o"][1, 1], b["I_o"][2, 2], b["I_o"][0, 1], b["I_o"][0, 2],
                             b["I_o"][1, 2]] for b in bodies]).reshape(nb, 6)

        geoms.sort(key=lambda g: g[0])
        for g in geoms:
            if g[2] not in _GEOM_CODE:
                raise NotImplementedError(
                    f"collision shape {g[2]!r} on link {g[6]!r}: only sphere/box/capsule/cylinder/half_space have "
                    "closed-form signed distance here (SURVEY.md §7 'hard parts')")
        # default collision filtering
        def body_parent(bi):
            return bodies[bi]["parent"] if bi >= 0 else None
        link_groups = {}
        for gname, members in self.filter_groups.items():
            for l in members:
                link_groups.setdefault(l, set()).add(gname)
        excl = set()
        for a, b in self.filter_excludes:
            excl.add((a, b))
            excl.add((b, a))
        adjacent = {(j.parent, j.child) for j in self.joints if j.type != "fixed" and j.parent != "world"}
        pairs = []
        for ia in range(len(geoms)):
            for ib in range(ia + 1, len(geoms)):
                ga, gb = geoms[ia], geoms[ib]
                ba, bb = ga[1], gb[1]
                if ba == bb:
                    continue  # same welded subgraph (merged body), or both anchored to the world
                # MultibodyPlant::ApplyDefaultCollisionFilters: the two bodies of a joint do not collide, except
                # when the joint's parent is the world itself or the joint is a free body's floating joint (a
                # ball may touch a palm that is welded to the world; a link may not touch the link it hinges on)
                if (ga[6], gb[6]) in adjacent or (gb[6], ga[6]) in adjacent:
                    continue
                ga_groups, gb_groups = link_groups.get(ga[6], set()), link_groups.get(gb[6], set())
                if any((x, y) in excl for x in ga_groups for y in gb_groups):
                    continue
                if ga[2] != "sphere" and gb[2] != "sphere":
                    raise NotImplementedError(
                        f"collision pair {ga[6]!r} ({ga[2]}) / {gb[6]!r} ({gb[2]}): closed-form signed distance needs "
                        "a sphere on one side (Drake falls back to FCL for the other pairs)")
                pairs.append((ia, ib))

        quat_starts = [q_start[k] for k, b in enumerate(bodies) if b["jtype"] == JOINT_QUAT_FLOATING]
        ng = len(geoms)
        return BakedModel(
            name=self.name, nbodies=nb, nq=nq, nv=nv,
            parent=np.array([b["parent"] for b in bodies], np.int32),
            joint_type=np.array([b["jtype"] for b in bodies], np.int32),
            q_start=np.array(q_start, np.int32), v_start=np.array(v_start, np.int32),
            X_PF=np.array([b["X_PF"].flat() for b in bodies]).reshape(nb, 12),
            R_MB=np.array([np.asarray(b["R_MB"]).reshape(-1) for b in bodies]).reshape(nb, 9),
            axis=np.array([b["axis"] for b in bodies], float).reshape(nb, 3),
            damping=damping, mass=mass, com=com, inertia=inertia, gravity=self.gravity.copy(),
            actuated=actuated, ngeoms=ng,
            geom_body=np.array([g[1] for g in geoms], np.int32),
            geom_type=np.array([_GEOM_CODE[g[2]] for g in geoms], np.int32),
            geom_dims=np.array([g[3] for g in geoms], float).reshape(ng, 3),
            X_BG=np.array([g[4].flat() for g in geoms]).reshape(ng, 12),
            npairs=len(pairs),
            pair_geomA=np.array([p[0] for p in pairs], np.int32),
            pair_geomB=np.array([p[1] for p in pairs], np.int32),
            body_names=[b["name"] for b in bodies], geom_names=[g[5] for g in geoms],
            quat_q_starts=quat_starts)


# ----------------------------------------------------------------------------
def pendulum_model():
    """The pendulum of the reference's closed-form tests (m=1, l=0.5, b=0.1, g=9.81;
    optimizer/test/trajectory_optimizer_test.cc:1109-1137).  Drake's Pendulum.urdf is not in
    the reference tree; the model is fully determined by those constants: a point mass at
    distance l below a revolute joint about +y, tau = m l^2 a + b v + m g l sin(q)."""
    m, l, b = 1.0, 0.5, 0.1
    return BakedModel(
        name="pendulum", nbodies=1, nq=1, nv=1,
        parent=np.array([-1], np.int32), joint_type=np.array([JOINT_REVOLUTE], np.int32),
        q_start=np.array([0], np.int32), v_start=np.array([0], np.int32),
        X_PF=X().flat().reshape(1, 12), R_MB=np.eye(3).reshape(1, 9),
        axis=np.array([[0.0, 1.0, 0.0]]), damping=np.array([b]), mass=np.array([m]),
        com=np.array([[0.0, 0.0, -l]]),
        inertia=np.array([[m * l * l, m * l * l, 0.0, 0.0, 0.0, 0.0]]),
        gravity=np.array([0.0, 0.0, -9.81]), actuated=np.array([1], np.int32), ngeoms=0,
        geom_body=np.zeros(0, np.int32), geom_type=np.zeros(0, np.int32),
        geom_dims=np.zeros((0, 3)), X_BG=np.zeros((0, 12)), npairs=0,
        pair_geomA=np.zeros(0, np.int32), pair_geomB=np.zeros(0, np.int32),
        body_names=["pendulum"], geom_names=[], quat_q_starts=[])


_MODEL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")


def load_model(name: str) -> BakedModel:
    """Loads a committed baked table (idto_b200/models/<name>.json)."""
    if name == "pendulum":
        return pendulum_model()
    return BakedModel.load(os.path.join(_MODEL_DIR, name + ".json"))


def bake_reference_models(reference_root="/root/reference", out_dir=_MODEL_DIR):
    """One-time bake of the reference's model files used by BASELINE.json's configs.
    Run in the build container (the reference tree does not exist on the GPU box)."""
    os.makedirs(out_dir, exist_ok=True)
    mdl = os.path.join(reference_root, "models")
    out = {}
    out["acrobot"] = ModelBuilder().add_urdf(os.path.join(mdl, "acrobot", "acrobot.urdf")).finalize()
    out["spinner"] = ModelBuilder().add_urdf(os.path.join(mdl, "spinner_friction.urdf")).finalize()
    out["spinner_sphere"] = ModelBuilder().add_urdf(os.path.join(mdl, "spinner_sphere.urdf")).finalize()
    out["spinner_capsule"] = ModelBuilder().add_urdf(os.path.join(mdl, "spinner_capsule.urdf")).finalize()
    b = ModelBuilder().add_urdf(os.path.join(mdl, "hopper.urdf"))
    out["hopper_no_ground"] = b.finalize()
    # examples/hopper/hopper.cc:44-50: ground Box(25,25,10) at z=-5 registered on the world body.
    b = ModelBuilder().add_urdf(os.path.join(mdl, "hopper.urdf"))
    b.register_collision_geometry("world", X(p=[0.0, 0.0, -5.0]), "box", [25.0, 25.0, 10.0], "ground")
    out["hopper"] = b.finalize()
    # the same ground as a HalfSpace (plant.RegisterCollisionGeometry(world_body, HalfSpace::MakePose(z, 0),
    # HalfSpace(), ...), the usual Drake ground): identical contact wherever the foot is over the box's top face
    b = ModelBuilder().add_urdf(os.path.join(mdl, "hopper.urdf"))
    b.register_collision_geometry("world", ModelBuilder.half_space_pose([0.0, 0.0, 1.0], [0.0, 0.0, 0.0]),
                                  "half_space", [], "ground")
    out["hopper_half_space"] = b.finalize()
    out["mini_cheetah"] = ModelBuilder().add_urdf(os.path.join(mdl, "mini_cheetah_with_ground.urdf")).finalize()
    # examples/allegro_hand/allegro_hand.cc:83-113: hand welded to the world with RPY(0, -pi/2, 0), free ball
    # (m = 0.05 kg, r = 0.06 m, solid sphere) with one collision sphere, registered after the hand's geometries
    b = ModelBuilder().add_sdf(os.path.join(mdl, "allegro_hand.sdf"))
    b.weld_frames("world", "hand_root", X(rpy_to_R(0.0, -math.pi / 2, 0.0), [0.0, 0.0, 0.0]))
    b.add_rigid_body("ball", 0.05, [0.0, 0.0, 0.0], 0.4 * 0.05 * 0.06 ** 2 * np.eye(3))
    b.register_collision_geometry("ball", X(), "sphere", [0.06], "ball_collision")
    out["allegro_hand"] = b.finalize()
    # --upside_down (allegro_hand.cc:94-97): the same plant with the gravity vector reversed
    b.gravity = np.array([0.0, 0.0, 9.81])
    out["allegro_hand_upside_down"] = b.finalize()
    for k, m in out.items():
        m.name = k
        m.save(os.path.join(out_dir, k + ".json"))
    return out


if __name__ == "__main__":
    for k, m in bake_reference_models().items():
        print(f"{k}: nbodies={m.nbodies} nq={m.nq} nv={m.nv} ngeoms={m.ngeoms} npairs={m.npairs} "
              f"unactuated={m.unactuated_dofs} bodies={m.body_names}")
