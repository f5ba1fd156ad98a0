"""CSGFoundry geometry arrays: container, builder and directory save/load.

The arrays are exactly the ones the reference persists and uploads (CSG/CSGFoundry.h:253-263,
save format CSG/CSGFoundry.cc:2768-2802):

    solid.npy (nsolid,3,4) int32   CSGSolid 48 B   label[16], numPrim, primOffset, type, pad, center_extent
    prim.npy  (nprim,4,4)  float32 CSGPrim  64 B   CSG/CSGPrim.h:72-118
    node.npy  (nnode,4,4)  float32 CSGNode  64 B   CSG/CSGNode.h:67-98
    tran.npy / itra.npy (ntran,4,4) float32 node transforms and their inverses (sysrap/sqat4.h)
    plan.npy  (nplan,4)    float32 planes of convex polyhedra
    inst.npy  (ninst,4,4)  float32 instance transforms, 4th column ints (sysrap/sqat4.h:345-407)

The builder follows CSGImport's conventions (CSG/CSGImport.cc:151-603): every leaf carries its
own 1-based transform index, boolean trees are complete binary trees in level order with the
node count in the root's subNum, differences are positivised into intersections with a
complemented right-hand side (sysrap/sn.h:2672-2698), and volumes that are not instanced are
flattened into solid 0 with world-frame node transforms and one identity instance.
"""
import os

import numpy as np

# typecodes : sysrap/OpticksCSG.h:21-62
CSG_ZERO = 0
CSG_UNION, CSG_INTERSECTION, CSG_DIFFERENCE = 1, 2, 3
CSG_CONTIGUOUS, CSG_DISCONTIGUOUS, CSG_OVERLAP = 11, 12, 13
CSG_SPHERE, CSG_ZSPHERE, CSG_CYLINDER, CSG_CONE, CSG_BOX3 = 101, 103, 105, 108, 110
CSG_CONVEXPOLYHEDRON, CSG_HYPERBOLOID, CSG_PHICUT, CSG_HALFSPACE = 112, 117, 121, 125

_UNBOUNDED = (CSG_PHICUT, CSG_HALFSPACE)


def translate(x, y, z):
    m = np.eye(4, dtype=np.float64)
    m[3, :3] = (x, y, z)          # row-vector convention: translation in elements 12..14
    return m


def rotate_z(deg):
    a = np.deg2rad(deg)
    m = np.eye(4, dtype=np.float64)
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
    return m


def rotate_x(deg):
    a = np.deg2rad(deg)
    m = np.eye(4, dtype=np.float64)
    m[1, 1], m[1, 2], m[2, 1], m[2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
    return m


def rotate_y(deg):
    a = np.deg2rad(deg)
    m = np.eye(4, dtype=np.float64)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a)
    return m


def scale(sx, sy, sz):
    return np.diag([sx, sy, sz, 1.0]).astype(np.float64)


class Leaf:
    """One CSG leaf in its own frame; `transform` (4x4, row-vector convention) places it in the
    frame of the solid."""

    def __init__(self, typecode, param, aabb, transform=None, complement=False, planes=None):
        self.typecode = typecode
        self.param = np.zeros(6, dtype=np.float32)
        self.param[:len(param)] = param
        self.aabb = None if aabb is None else np.asarray(aabb, dtype=np.float64)   # lo.xyz hi.xyz local frame
        self.transform = np.eye(4) if transform is None else np.asarray(transform, dtype=np.float64)
        self.complement = complement
        self.planes = planes

    def placed(self, m):
        """copy with an additional transform applied after the leaf's own"""
        return Leaf(self.typecode, self.param, self.aabb, self.transform @ m, self.complement, self.planes)

    def complemented(self):
        return Leaf(self.typecode, self.param, self.aabb, self.transform, not self.complement, self.planes)


def sphere(r):
    return Leaf(CSG_SPHERE, [0, 0, 0, r], [-r, -r, -r, r, r, r])


def zsphere(r, z1, z2):
    return Leaf(CSG_ZSPHERE, [0, 0, 0, r, z1, z2], [-r, -r, z1, r, r, z2])


def box3(fx, fy, fz):
    return Leaf(CSG_BOX3, [fx, fy, fz, 0], [-fx / 2, -fy / 2, -fz / 2, fx / 2, fy / 2, fz / 2])


def cylinder(r, z1, z2):
    return Leaf(CSG_CYLINDER, [0, 0, 0, r, z1, z2], [-r, -r, z1, r, r, z2])


def cone(r1, z1, r2, z2):
    rm = max(r1, r2)
    return Leaf(CSG_CONE, [r1, z1, r2, z2], [-rm, -rm, z1, rm, rm, z2])


def hyperboloid(r0, zf, z1, z2):
    rm = r0 * np.sqrt(max((z1 / zf) ** 2, (z2 / zf) ** 2) + 1.0)
    return Leaf(CSG_HYPERBOLOID, [r0, zf, z1, z2], [-rm, -rm, z1, rm, rm, z2])


def phicut(phi0_deg, phi1_deg):
    a0, a1 = np.deg2rad(phi0_deg), np.deg2rad(phi1_deg)
    return Leaf(CSG_PHICUT, [np.cos(a0), np.sin(a0), np.cos(a1), np.sin(a1)], None)


def halfspace(nx, ny, nz, w):
    return Leaf(CSG_HALFSPACE, [nx, ny, nz, w], None)


def convexpolyhedron(planes, aabb):
    """planes: (n,4) outward normal + distance; aabb must be given (CSG::ExpectExternalBBox)"""
    return Leaf(CSG_CONVEXPOLYHEDRON, [0, 0, 0, 0], aabb, planes=np.asarray(planes, dtype=np.float32))


class Op:
    """boolean operator over two subtrees (Leaf or Op)"""

    def __init__(self, typecode, left, right):
        self.typecode = typecode
        self.left = left
        self.right = right


def union(a, b):
    return Op(CSG_UNION, a, b)


def intersection(a, b):
    return Op(CSG_INTERSECTION, a, b)


def difference(a, b):
    return Op(CSG_DIFFERENCE, a, b)


class ListNode:
    """multi-union / multi-intersection of leaves (CSG_CONTIGUOUS, CSG_DISCONTIGUOUS, CSG_OVERLAP)"""

    def __init__(self, typecode, subs):
        self.typecode = typecode
        self.subs = list(subs)


def _positivize(t, negate=False):
    """sn::positivize (sysrap/sn.h:2672-2698): A - B -> A * !B, De Morgan for negated subtrees."""
    if isinstance(t, Leaf):
        return t.complemented() if negate else t
    if isinstance(t, ListNode):
        assert not negate, "complemented list nodes are not supported"
        return t
    op = t.typecode
    if op == CSG_DIFFERENCE:
        if negate:      # !(A - B) = !A + B
            return Op(CSG_UNION, _positivize(t.left, True), _positivize(t.right, False))
        return Op(CSG_INTERSECTION, _positivize(t.left, False), _positivize(t.right, True))
    if op == CSG_UNION:
        if negate:
            return Op(CSG_INTERSECTION, _positivize(t.left, True), _positivize(t.right, True))
        return Op(CSG_UNION, _positivize(t.left), _positivize(t.right))
    if op == CSG_INTERSECTION:
        if negate:
            return Op(CSG_UNION, _positivize(t.left, True), _positivize(t.right, True))
        return Op(CSG_INTERSECTION, _positivize(t.left), _positivize(t.right))
    raise ValueError(op)


def _height(t):
    if isinstance(t, Op):
        return 1 + max(_height(t.left), _height(t.right))
    return 0


def _level_order(t, height):
    """complete binary tree of the given height as a 1-based level-order list, None = CSG_ZERO"""
    slots = [None] * ((1 << (height + 1)) - 1)

    def put(node, idx):
        slots[idx - 1] = node
        if isinstance(node, Op):
            put(node.left, 2 * idx)
            put(node.right, 2 * idx + 1)

    put(t, 1)
    return slots


def _transform_aabb(aabb, m):
    lo, hi = aabb[:3], aabb[3:]
    c = np.array([[(hi if (k >> a) & 1 else lo)[a] for a in range(3)] for k in range(8)])
    w = c @ m[:3, :3] + m[3, :3]
    return np.concatenate([w.min(0), w.max(0)])


def affine_inverse(m):
    """Inverse of a 4x4 affine transform in the row-vector convention (linear part in rows 0-2, translation in row 3),
    by cofactors in double precision like glm::inverse (sysrap/stran.h Tran<double>): a pure translation - and any
    linear part made of 0 / +-1 entries - inverts EXACTLY, where an LU factorisation (np.linalg.inv) leaves 1e-17-sized
    entries off the diagonal that a CSGFoundry written by the reference does not have."""
    m = np.asarray(m, dtype=np.float64)
    a = m[:3, :3]
    c = np.empty((3, 3), dtype=np.float64)
    c[0, 0] = a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]
    c[0, 1] = a[0, 2] * a[2, 1] - a[0, 1] * a[2, 2]
    c[0, 2] = a[0, 1] * a[1, 2] - a[0, 2] * a[1, 1]
    c[1, 0] = a[1, 2] * a[2, 0] - a[1, 0] * a[2, 2]
    c[1, 1] = a[0, 0] * a[2, 2] - a[0, 2] * a[2, 0]
    c[1, 2] = a[0, 2] * a[1, 0] - a[0, 0] * a[1, 2]
    c[2, 0] = a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]
    c[2, 1] = a[0, 1] * a[2, 0] - a[0, 0] * a[2, 1]
    c[2, 2] = a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]
    det = a[0, 0] * c[0, 0] + a[0, 1] * c[1, 0] + a[0, 2] * c[2, 0]
    if det == 0.0 or not np.isfinite(det):
        raise ValueError("singular transform")
    inv = np.zeros((4, 4), dtype=np.float64)
    inv[:3, :3] = c / det
    inv[3, :3] = -(m[3, 0] * inv[0, :3] + m[3, 1] * inv[1, :3] + m[3, 2] * inv[2, :3])
    inv[3, 3] = 1.0
    inv[np.abs(inv) == 0.0] = 0.0          # no negative zeros
    return inv


class Foundry:
    """Accumulates solids / prims / nodes and emits the CSGFoundry arrays."""

    def __init__(self):
        self.solids = []     # (label, numPrim, primOffset)
        self.prims = []      # dict
        self.nodes = []      # (16,) float32 rows
        self.trans = []      # 4x4 float64
        self.planes = []     # (4,)
        self.insts = []      # (4x4 float64, ins_idx, gas_idx, sensor_id+1, sensor_index)
        self.meshnames = []
        self._open = None

    # ---- building --------------------------------------------------------------------------
    def begin_solid(self, label):
        assert self._open is None
        self._open = (label, len(self.prims))

    def end_solid(self):
        label, off = self._open
        self.solids.append((label, len(self.prims) - off, off))
        self._open = None
        return len(self.solids) - 1

    def _add_node(self, typecode, boundary, param=None, aabb=None, tran_idx=0, complement=False, sub=None):
        row = np.zeros(16, dtype=np.float32)
        u = row.view(np.uint32)
        if param is not None:
            row[0:6] = param
        if sub is not None:
            u[0], u[1] = sub
        u[6] = boundary
        u[7] = len(self.nodes)
        if aabb is not None:
            row[8:14] = aabb
        u[14] = typecode
        u[15] = (tran_idx & 0x7fffffff) | (0x80000000 if complement else 0)
        self.nodes.append(row)
        return len(self.nodes) - 1

    def _add_leaf(self, leaf, boundary, frame):
        m = leaf.transform @ frame
        self.trans.append(m)
        tran_idx = len(self.trans)                      # 1-based
        aabb = None
        if leaf.aabb is not None:
            aabb = _transform_aabb(leaf.aabb, m)
        param = leaf.param.copy()
        if leaf.typecode == CSG_CONVEXPOLYHEDRON:
            pu = param.view(np.uint32)
            pu[0] = len(self.planes)
            pu[1] = len(leaf.planes)
            self.planes.extend(list(leaf.planes))
        self._add_node(leaf.typecode, boundary, param, aabb, tran_idx, leaf.complement)
        return aabb if not leaf.complement else None

    def add_prim(self, shape, boundary, frame=None, mesh_idx=0, name=None):
        """One CSGPrim from a Leaf, an Op tree or a ListNode, placed by `frame` (4x4)."""
        assert self._open is not None, "begin_solid first"
        frame = np.eye(4) if frame is None else np.asarray(frame, dtype=np.float64)
        node_offset = len(self.nodes)
        tran_offset = len(self.trans)
        plan_offset = len(self.planes)
        boxes = []
        if isinstance(shape, Leaf):
            boxes.append(self._add_leaf(shape, boundary, frame))
        elif isinstance(shape, ListNode):
            self._add_node(shape.typecode, boundary, sub=(len(shape.subs), 1))
            for s in shape.subs:
                boxes.append(self._add_leaf(s, boundary, frame))
        else:
            tree = _positivize(shape)
            h = _height(tree)
            slots = _level_order(tree, h)
            bn = len(slots)
            lists = []
            sub_offset = bn
            for i, nd in enumerate(slots):
                if nd is None:
                    self._add_node(CSG_ZERO, boundary)
                elif isinstance(nd, Op):
                    self._add_node(nd.typecode, boundary, sub=(bn, 0) if i == 0 else None)
                elif isinstance(nd, ListNode):
                    self._add_node(nd.typecode, boundary, sub=(len(nd.subs), sub_offset))
                    sub_offset += len(nd.subs)
                    lists.append(nd)
                else:
                    boxes.append(self._add_leaf(nd, boundary, frame))
            for nd in lists:
                for s in nd.subs:
                    boxes.append(self._add_leaf(s, boundary, frame))
        boxes = [b for b in boxes if b is not None]
        assert boxes, "prim needs at least one bounded, uncomplemented leaf for its AABB"
        bb = np.concatenate([np.min([b[:3] for b in boxes], 0), np.max([b[3:] for b in boxes], 0)])
        self.prims.append(dict(num_node=len(self.nodes) - node_offset, node_offset=node_offset,
                               tran_offset=tran_offset, plan_offset=plan_offset, mesh_idx=mesh_idx,
                               repeat_idx=len(self.solids), prim_idx=len(self.prims) - self._open[1], aabb=bb))
        self.meshnames.append(name or "prim%d" % (len(self.prims) - 1))
        return len(self.prims) - 1

    def add_instance(self, transform, gas_idx, sensor_identifier=-1, sensor_index=-1):
        """sensor_identifier -1 = not a sensor; stored +1 like CSGFoundry::addInstanceVector
        (CSG/CSGFoundry.cc:2156-2169)"""
        self.insts.append((np.asarray(transform, dtype=np.float64), len(self.insts), gas_idx, sensor_identifier + 1, sensor_index))

    # ---- arrays ----------------------------------------------------------------------------
    def arrays(self):
        assert self._open is None
        ns, npr, nn = len(self.solids), len(self.prims), len(self.nodes)
        solid = np.zeros((ns, 3, 4), dtype=np.int32)
        for i, (label, num, off) in enumerate(self.solids):
            lab = label.encode()[:15].ljust(16, b"\0")
            solid[i].reshape(-1).view(np.uint8)[:16] = np.frombuffer(lab, dtype=np.uint8)
            solid[i, 1, 0], solid[i, 1, 1], solid[i, 1, 2] = num, off, 0
            bbs = np.array([p["aabb"] for p in self.prims[off:off + num]])
            lo, hi = bbs[:, :3].min(0), bbs[:, 3:].max(0)
            ce = np.concatenate([(lo + hi) / 2, [np.max(hi - lo) / 2]]).astype(np.float32)
            solid[i, 2].view(np.float32)[:] = ce
        prim = np.zeros((npr, 4, 4), dtype=np.float32)
        pi = prim.view(np.int32)
        for i, p in enumerate(self.prims):
            pi[i, 0] = (p["num_node"], p["node_offset"], p["tran_offset"], p["plan_offset"])
            pi[i, 1] = (i, p["mesh_idx"], p["repeat_idx"], p["prim_idx"])
            prim[i].reshape(-1)[8:14] = p["aabb"]
            pi[i, 3, 3] = i                      # globalPrimIdx
        node = np.array(self.nodes, dtype=np.float32).reshape(nn, 4, 4)
        tran = np.array(self.trans, dtype=np.float64).reshape(-1, 4, 4)
        itra = np.array([affine_inverse(t) for t in tran]).reshape(-1, 4, 4)
        plan = np.array(self.planes, dtype=np.float32).reshape(-1, 4)
        insts = self.insts or [(np.eye(4), 0, 0, 0, -1)]
        inst = np.zeros((len(insts), 4, 4), dtype=np.float32)
        ii = inst.view(np.int32)
        for i, (m, ins_idx, gas_idx, sid1, sidx) in enumerate(insts):
            inst[i] = m
            ii[i, 0, 3], ii[i, 1, 3], ii[i, 2, 3], ii[i, 3, 3] = ins_idx, gas_idx, sid1, sidx
        return dict(solid=solid, prim=prim, node=node, tran=tran.astype(np.float32), itra=itra.astype(np.float32),
                    plan=plan, inst=inst, meshname=list(self.meshnames))


def save_foundry(arrays, folder):
    """CSGFoundry::save_ layout (CSG/CSGFoundry.cc:2768-2802)"""
    os.makedirs(folder, exist_ok=True)
    for k in ("solid", "prim", "node", "tran", "itra", "inst"):
        np.save(os.path.join(folder, k + ".npy"), arrays[k])
    if len(arrays["plan"]):
        np.save(os.path.join(folder, "plan.npy"), arrays["plan"])
    with open(os.path.join(folder, "meshname.txt"), "w") as f:
        f.write("\n".join(arrays.get("meshname", [])) + "\n")


def load_foundry(folder):
    out = {}
    for k in ("solid", "prim", "node", "tran", "itra", "inst"):
        out[k] = np.load(os.path.join(folder, k + ".npy"))
    p = os.path.join(folder, "plan.npy")
    out["plan"] = np.load(p) if os.path.exists(p) else np.zeros((0, 4), dtype=np.float32)
    m = os.path.join(folder, "meshname.txt")
    out["meshname"] = open(m).read().split("\n")[:-1] if os.path.exists(m) else []
    return out


def save_geometry(geom, folder):
    """Persist foundry + tables the way a reference geometry directory is laid out:
    <folder>/CSGFoundry/*.npy and <folder>/CSGFoundry/SSim/stree/standard/{bnd,optical,icdf}.npy"""
    fd = os.path.join(folder, "CSGFoundry")
    save_foundry(geom["foundry"], fd)
    ss = os.path.join(fd, "SSim", "stree", "standard")
    os.makedirs(ss, exist_ok=True)
    np.save(os.path.join(ss, "bnd.npy"), np.ascontiguousarray(geom["bnd"], dtype=np.float32))
    np.save(os.path.join(ss, "optical.npy"), np.ascontiguousarray(geom["optical"], dtype=np.int32))
    if geom.get("icdf") is not None:
        np.save(os.path.join(ss, "icdf.npy"), np.ascontiguousarray(geom["icdf"], dtype=np.float32))
    with open(os.path.join(ss, "bnd_names.txt"), "w") as f:
        f.write("\n".join(geom.get("bnd_names", [])) + "\n")


def load_geometry(folder):
    fd = os.path.join(folder, "CSGFoundry")
    ss = os.path.join(fd, "SSim", "stree", "standard")
    out = dict(foundry=load_foundry(fd), bnd=np.load(os.path.join(ss, "bnd.npy")), optical=np.load(os.path.join(ss, "optical.npy")))
    p = os.path.join(ss, "icdf.npy")
    out["icdf"] = np.load(p) if os.path.exists(p) else None
    n = os.path.join(ss, "bnd_names.txt")
    out["bnd_names"] = open(n).read().split("\n")[:-1] if os.path.exists(n) else []
    return out
