"""A small SPIR-V interpreter: executes the reference's SHIPPED compute shaders (shaders/*.comp.spv of Thefefe/orbit)
on the CPU so that the oracle can be pinned against the reference's own GPU programs.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): used by tests/golden/make_spirv_golden.py, in this container,
to produce the fixtures tests/golden/spirv_*.json from /root/reference/shaders/*.spv. Nothing in orbit_b200/ imports it.

What is interpreted faithfully: the module's types, constants, specialisation constants, explicit buffer layouts
(Offset / ArrayStride / MatrixStride), bindless descriptor indexing, structured control flow with OpPhi, integer and
floating-point arithmetic (every FP instruction individually rounded to binary32, OpExtInst Fma fused), conversions,
bit operations, atomics, subgroup ballot / elect (subgroup = 32 consecutive invocations; an operation sees the
invocations waiting at the same instruction), image fetch / store / explicit-LOD sampling.

What SPIR-V / Vulkan leave to the implementation and this VM pins the same way as DESIGN.md §3:
  * OpDot, OpMatrixTimesVector, OpMatrixTimesMatrix, Length, Distance: products summed left to right, no contraction;
  * Log2: the contract's orbit_log2f (passed in as `log2f`);
  * the LINEAR + MIN-reduction sampler: level = nearest(lod) (ceil(lod + 0.5) - 1, clamped), footprint
    i0 = floor(u*w - 0.5), i1 = i0 + 1, clamp-to-edge, minimum of the four texels.
"""
import math
import struct

import numpy as np

from spv_parse import Module

F32 = np.float32


def f32(x):
    return F32(x)


def fma32(a, b, c):
    """Correctly rounded binary32 fused multiply-add."""
    a, b, c = float(a), float(b), float(c)
    p = a * b                        # exact in binary64 (24 x 24 bits)
    s = p + c
    if math.isnan(s) or math.isinf(s):
        return F32(s)
    # error of the binary64 addition (TwoSum)
    bb = s - p
    err = (p - (s - bb)) + (c - bb)
    r = F32(s)
    if err != 0.0:
        # s was rounded in binary64: only matters if s sits exactly halfway between two binary32 values
        lo = float(np.nextafter(r, F32(-np.inf))); hi = float(np.nextafter(r, F32(np.inf)))
        rf = float(r)
        if s == (rf + lo) / 2.0 or s == (rf + hi) / 2.0 or rf == s:
            # decide with the exact sign of the residual
            from fractions import Fraction
            exact = Fraction(a) * Fraction(b) + Fraction(c)
            cands = sorted({lo, rf, hi})
            best = min(cands, key=lambda v: (abs(Fraction(v) - exact), int(struct.unpack("<I", struct.pack("<f", v))[0] & 1)))
            return F32(best)
    return r


class Buffer:
    def __init__(self, data):
        self.data = np.frombuffer(bytearray(bytes(data)), dtype=np.uint8).copy() if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1).copy()

    def read(self, off, n):
        return bytes(self.data[off:off + n])

    def write(self, off, b):
        self.data[off:off + len(b)] = np.frombuffer(b, dtype=np.uint8)


class Image:
    """levels: list of np arrays [h, w] (float32) or [h, w, c] / [d, h, w, c] (uint32) for storage images."""

    def __init__(self, levels):
        self.levels = levels


class Sampler:
    def __init__(self, reduce_min=True):
        self.reduce_min = reduce_min


class Ptr:
    __slots__ = ("kind", "base", "off", "type", "mstride", "path")

    def __init__(self, kind, base, off, type_, mstride=0, path=None):
        self.kind, self.base, self.off, self.type, self.mstride, self.path = kind, base, off, type_, mstride, path


class Wait(Exception):
    pass


class VM:
    def __init__(self, path, spec=None, log2f=None, contract_fma=False):
        # contract_fma=True: a*b+c chains in OpDot / OpMatrixTimesVector are fused the way an optimising driver compiler
        # typically does (sensitivity experiments only; the fixtures use the contract's unfused, left-to-right sums)
        self.contract_fma = contract_fma
        self.m = m = Module(path)
        self.spec = spec or {}
        self.log2f = log2f or (lambda x: F32(math.log2(float(x))) if x > 0 else F32(-np.inf))
        self.types, self.consts, self.globals = {}, {}, {}
        self.resources = {}            # (set, binding) -> {index: Buffer | Image | Sampler}
        self.push = Buffer(b"\0" * 256)
        self.trace = False
        self.emitted = []              # task shaders: (workgroup id, [(group counts, payload)]) per workgroup
        self.blocks, self.block_order = {}, []
        self._decode()

    # ---- module decoding -------------------------------------------------------------------------------
    def _decode(self):
        m = self.m
        cur = None
        for inst in m.insts:
            op, a = inst.op, inst.words
            if op == 19: self.types[a[0]] = ("void",)
            elif op == 20: self.types[a[0]] = ("bool",)
            elif op == 21: self.types[a[0]] = ("int", a[1], a[2])
            elif op == 22: self.types[a[0]] = ("float", a[1])
            elif op == 23: self.types[a[0]] = ("vec", a[1], a[2])
            elif op == 24: self.types[a[0]] = ("mat", a[1], a[2])
            elif op == 25: self.types[a[0]] = ("image", a[1], a[2], a[6] if len(a) > 6 else 0)   # sampled type, dim, sampled flag
            elif op == 26: self.types[a[0]] = ("sampler",)
            elif op == 27: self.types[a[0]] = ("sampled_image", a[1])
            elif op == 28: self.types[a[0]] = ("array", a[1], a[2])
            elif op == 29: self.types[a[0]] = ("rtarray", a[1])
            elif op == 30: self.types[a[0]] = ("struct", list(a[1:]))
            elif op == 32: self.types[a[0]] = ("ptr", a[1], a[2])
            elif op == 33: self.types[a[0]] = ("fn",)
            elif op in (41, 42): self.consts[a[1]] = (op == 41)
            elif op in (43, 50):
                v = self._scalar_const(a[0], a[2:])
                if op == 50:
                    sid = m.decor.get(a[1], {}).get(1)
                    if sid is not None and sid[0] in self.spec:
                        v = self.spec[sid[0]]
                        if self.types[a[0]][0] == "float": v = F32(v)
                self.consts[a[1]] = v
            elif op == 48 or op == 49:   # SpecConstantTrue / False
                sid = m.decor.get(a[1], {}).get(1)
                self.consts[a[1]] = bool(self.spec.get(sid[0], op == 48)) if sid else (op == 48)
            elif op in (44, 51): self.consts[a[1]] = [self.consts[x] for x in a[2:]]
            elif op == 46: self.consts[a[1]] = self._null(a[0])
            elif op == 1: self.consts[a[1]] = self._null(a[0])
            elif op == 52: self.consts[a[1]] = self._spec_op(a)
            elif op == 59 and cur is None:
                self.globals[a[1]] = inst
            elif op == 248:
                cur = a[0]; self.blocks[cur] = []; self.block_order.append(cur)
            elif op == 56:
                cur = None
            elif cur is not None:
                self.blocks[cur].append(inst)

    def _scalar_const(self, t, words):
        ty = self.types[t]
        if ty[0] == "float":
            assert ty[1] == 32
            return F32(struct.unpack("<f", struct.pack("<I", words[0]))[0])
        if ty[0] == "int":
            return words[0] & ((1 << ty[1]) - 1)
        raise NotImplementedError(ty)

    def _null(self, t):
        ty = self.types[t]
        k = ty[0]
        if k == "bool": return False
        if k == "int": return 0
        if k == "float": return F32(0)
        if k == "vec": return [self._null(ty[1]) for _ in range(ty[2])]
        if k == "mat": return [self._null(ty[1]) for _ in range(ty[2])]
        if k == "array": return [self._null(ty[1]) for _ in range(self.consts[ty[2]])]
        if k == "struct": return [self._null(x) for x in ty[1]]
        return None

    def _spec_op(self, a):
        t, op = a[0], a[2]
        args = [self.consts[x] for x in a[3:]]
        w = self.types[t][1] if self.types[t][0] == "int" else 32
        mask = (1 << w) - 1
        if op == 128: return (args[0] + args[1]) & mask
        if op == 130: return (args[0] - args[1]) & mask
        if op == 132: return (args[0] * args[1]) & mask
        if op == 134: return (args[0] // args[1]) & mask if args[1] else 0
        if op == 137: return (args[0] % args[1]) & mask if args[1] else 0
        if op == 194: return (args[0] >> args[1]) & mask
        if op == 196: return (args[0] << args[1]) & mask
        if op == 113: return args[0] & mask
        raise NotImplementedError("SpecConstantOp %d" % op)

    # ---- layout ------------------------------------------------------------------------------------------
    def size_of(self, t):
        ty = self.types[t]
        k = ty[0]
        if k in ("int", "float"): return ty[1] // 8
        if k == "vec": return self.size_of(ty[1]) * ty[2]
        raise NotImplementedError(ty)

    def load_mem(self, buf, off, t, mstride=0):
        ty = self.types[t]
        k = ty[0]
        if k == "int":
            n = ty[1] // 8
            return int.from_bytes(buf.read(off, n), "little")
        if k == "float":
            return F32(struct.unpack("<f", buf.read(off, 4))[0])
        if k == "vec":
            s = self.size_of(ty[1])
            return [self.load_mem(buf, off + i * s, ty[1]) for i in range(ty[2])]
        if k == "mat":
            assert mstride, "matrix without MatrixStride"
            return [self.load_mem(buf, off + c * mstride, ty[1]) for c in range(ty[2])]
        if k == "array":
            st = self.m.decor[t][6][0]
            return [self.load_mem(buf, off + i * st, ty[1], mstride) for i in range(self.consts[ty[2]])]
        if k == "struct":
            out = []
            for i, mt in enumerate(ty[1]):
                d = self.m.member_decor.get((t, i), {})
                assert 4 not in d, "RowMajor not supported"
                out.append(self.load_mem(buf, off + d[35][0], mt, d.get(7, [0])[0]))
            return out
        raise NotImplementedError(ty)

    def store_mem(self, buf, off, t, v, mstride=0):
        ty = self.types[t]
        k = ty[0]
        if k == "int":
            buf.write(off, int(v).to_bytes(ty[1] // 8, "little"))
        elif k == "float":
            buf.write(off, struct.pack("<f", float(v)))
        elif k == "vec":
            s = self.size_of(ty[1])
            for i in range(ty[2]): self.store_mem(buf, off + i * s, ty[1], v[i])
        elif k == "mat":
            for c in range(ty[2]): self.store_mem(buf, off + c * mstride, ty[1], v[c])
        elif k == "array":
            st = self.m.decor[t][6][0]
            for i in range(self.consts[ty[2]]): self.store_mem(buf, off + i * st, ty[1], v[i], mstride)
        elif k == "struct":
            for i, mt in enumerate(ty[1]):
                d = self.m.member_decor.get((t, i), {})
                self.store_mem(buf, off + d[35][0], mt, v[i], d.get(7, [0])[0])
        else:
            raise NotImplementedError(ty)

    # ---- dispatch ----------------------------------------------------------------------------------------
    def dispatch(self, groups, local_size, subgroup=32):
        """Runs groups[0] x groups[1] x groups[2] workgroups of local_size invocations each."""
        lx, ly, lz = local_size
        for gz in range(groups[2]):
            for gy in range(groups[1]):
                for gx in range(groups[0]):
                    invs = []
                    shared = {}
                    emitted = []
                    self.emitted.append(((gx, gy, gz), emitted))
                    for z in range(lz):
                        for y in range(ly):
                            for x in range(lx):
                                li = x + lx * (y + ly * z)
                                b = {"local": [x, y, z], "group": [gx, gy, gz], "global": [gx * lx + x, gy * ly + y, gz * lz + z],
                                     "num_groups": list(groups), "local_index": li, "subgroup_size": subgroup,
                                     "subgroup_inv": li % subgroup, "subgroup_id": li // subgroup, "wg_size": [lx, ly, lz],
                                     "shared": shared, "emit": (lambda counts, payload, e=emitted: e.append((counts, payload)))}
                                invs.append(self._run(b))
                    self._schedule(invs, subgroup)

    def _schedule(self, invs, subgroup):
        """Runs one workgroup. Every invocation runs until it finishes or reaches a scheduling point: a subgroup operation,
        a barrier, or an atomic. The pending point with the EARLIEST instruction is served first, for all invocations standing
        at it (subgroup operations: per subgroup), so the workgroup advances through the program front to back — one legal
        interleaving, and the one in which 'every invocation has appended before anyone reads the total' holds for shaders
        that rely on it without a barrier (active_cluster_compaction.comp does)."""
        state = [None] * len(invs)           # pending (kind, instruction index, payload)
        alive = [True] * len(invs)

        def advance(i, send=None):
            try:
                state[i] = invs[i].send(send)
            except StopIteration:
                alive[i] = False; state[i] = None
        for i in range(len(invs)):
            advance(i)
        while any(alive):
            live = [i for i in range(len(invs)) if alive[i]]
            key = min(state[i][1] for i in live)
            group = [i for i in live if state[i][1] == key]
            kind = state[group[0]][0]
            if kind == "barrier":
                if len(group) != len(live):
                    raise RuntimeError("workgroup barrier not reached by every live invocation")
                for i in group: advance(i)
            elif kind == "sync":
                for i in group: advance(i)
            elif kind in ("ballot", "elect"):
                for sg in range(0, len(invs), subgroup):
                    lanes = [i for i in group if sg <= i < sg + subgroup]
                    if not lanes:
                        continue
                    if kind == "ballot":
                        mask = 0
                        for i in lanes:
                            if state[i][2]: mask |= 1 << (i - sg)
                        res = [mask & 0xFFFFFFFF, 0, 0, 0]
                        for i in lanes: advance(i, list(res))
                    else:
                        first = min(lanes)
                        for i in lanes: advance(i, i == first)
            else:
                raise NotImplementedError(kind)

    # ---- one invocation ------------------------------------------------------------------------------------
    def _run(self, builtins):
        m, types, consts = self.m, self.types, self.consts
        vals = {}
        mem_vars = {}

        def V(i):
            return vals[i] if i in vals else consts[i]

        def var_pointer(vid):
            inst = self.globals.get(vid)
            if inst is None:
                return vals[vid]
            pt = types[inst.words[0]]
            sc = pt[1]
            dec = m.decor.get(vid, {})
            if 11 in dec:    # BuiltIn
                bi = dec[11][0]
                name = {24: "num_groups", 25: "wg_size", 26: "group", 27: "local", 28: "global", 29: "local_index", 36: "subgroup_size",
                        41: "subgroup_inv", 40: "subgroup_id"}[bi]
                return Ptr("var", [builtins[name]], 0, pt[2], path=[0])
            if sc == 9:      # PushConstant
                return Ptr("mem", self.push, 0, pt[2])
            if sc in (12, 2, 0):   # StorageBuffer / Uniform / UniformConstant: descriptor (array)
                return Ptr("desc", (dec[34][0], dec[33][0]), 0, pt[2])
            if sc in (4, 5402):   # Workgroup / TaskPayloadWorkgroupEXT: one instance per workgroup
                shared = builtins["shared"]
                if vid not in shared:
                    shared[vid] = [self._null(pt[2])]
                return Ptr("var", shared[vid], 0, pt[2], path=[0])
            if sc in (6, 7):  # Private / Function (module-scope private)
                if vid not in mem_vars:
                    mem_vars[vid] = [self._null(pt[2])]
                return Ptr("var", mem_vars[vid], 0, pt[2], path=[0])
            raise NotImplementedError("storage class %d" % sc)

        def access(base, idxs):
            p = base if isinstance(base, Ptr) else var_pointer(base)
            kind, b, off, t, ms, path = p.kind, p.base, p.off, p.type, p.mstride, list(p.path or [])
            for ix in idxs:
                ty = types[t]
                k = ty[0]
                if kind == "desc":
                    if k in ("array", "rtarray"):
                        res = self.resources[b][ix]
                        t = ty[1]
                        if isinstance(res, Buffer):
                            kind, b, off = "mem", res, 0
                        else:
                            kind, b = "res", res
                        continue
                    raise NotImplementedError("descriptor access")
                if k == "struct":
                    if kind == "mem":
                        d = m.member_decor[(t, ix)]
                        off += d[35][0]; ms = d.get(7, [0])[0]
                    else:
                        path.append(ix)
                    t = ty[1][ix]
                elif k in ("array", "rtarray"):
                    if kind == "mem": off += ix * m.decor[t][6][0]
                    else: path.append(ix)
                    t = ty[1]
                elif k == "mat":
                    if kind == "mem": off += ix * ms
                    else: path.append(ix)
                    t = ty[1]
                elif k == "vec":
                    if kind == "mem": off += ix * self.size_of(ty[1])
                    else: path.append(ix)
                    t = ty[1]
                else:
                    raise NotImplementedError(ty)
            return Ptr(kind, b, off, t, ms, path)

        def load(p):
            if not isinstance(p, Ptr): p = var_pointer(p)
            if p.kind == "mem": return self.load_mem(p.base, p.off, p.type, p.mstride)
            if p.kind == "val": return p.base
            if p.kind == "res": return p.base
            if p.kind == "desc":
                # a non-arrayed descriptor
                return self.resources[p.base][0]
            c = p.base
            for ix in p.path: c = c[ix]
            return c

        def store(p, v):
            if not isinstance(p, Ptr): p = var_pointer(p)
            if p.kind == "mem":
                self.store_mem(p.base, p.off, p.type, v, p.mstride); return
            c = p.base
            for ix in p.path[:-1]: c = c[ix]
            c[p.path[-1]] = v

        def width(t):
            ty = types[t]
            if ty[0] == "vec": ty = types[ty[1]]
            return ty[1] if ty[0] in ("int", "float") else 32

        def vmap(f, *xs):
            if isinstance(xs[0], list): return [f(*[x[i] if isinstance(x, list) else x for x in xs]) for i in range(len(xs[0]))]
            return f(*xs)

        def signed(v, w):
            return v - (1 << w) if v >> (w - 1) else v

        fuse = self.contract_fma

        def dot(a, b):
            acc = a[0] * b[0]
            for i in range(1, len(a)): acc = fma32(a[i], b[i], acc) if fuse else F32(acc + a[i] * b[i])
            return acc

        def mat_vec(M, v):
            out = []
            for r in range(len(M[0])):
                acc = M[0][r] * v[0]
                for c in range(1, len(M)): acc = fma32(M[c][r], v[c], acc) if fuse else F32(acc + M[c][r] * v[c])
                out.append(acc)
            return out

        label, prev = self.block_order[0], None
        with np.errstate(all="ignore"):
            while True:
                insts = self.blocks[label]
                # phis first (parallel copy)
                newv = {}
                k = 0
                while k < len(insts) and insts[k].op in (245, 8, 317):
                    if insts[k].op == 245:
                        a = insts[k].words
                        for j in range(2, len(a), 2):
                            if a[j + 1] == prev:
                                newv[a[1]] = V(a[j]); break
                        else:
                            raise RuntimeError("phi without matching predecessor")
                    k += 1
                vals.update(newv)
                nxt = None
                for inst in insts[k:]:
                    op, a = inst.op, inst.words
                    if op in (8, 317, 246, 247): continue
                    if self.trace and inst.result is not None and inst.result in vals: pass
                    if self.trace: print("exec", op, list(a))
                    if op == 61: vals[a[1]] = load(a[2] if a[2] not in vals else vals[a[2]])
                    elif op == 62: store(a[0] if a[0] not in vals else vals[a[0]], V(a[1]))
                    elif op in (65, 66): vals[a[1]] = access(a[2] if a[2] not in vals else vals[a[2]], [V(x) for x in a[3:]])
                    elif op == 59:
                        cell = [V(a[3]) if len(a) > 3 else self._null(types[a[0]][2])]
                        vals[a[1]] = Ptr("var", cell, 0, types[a[0]][2], path=[0])
                    elif op == 81:
                        c = V(a[2])
                        for ix in a[3:]: c = c[ix]
                        vals[a[1]] = c
                    elif op == 80: vals[a[1]] = self._construct(a[0], [V(x) for x in a[2:]])
                    elif op == 82:
                        obj = self._copy(V(a[3])); c = obj
                        for ix in a[4:-1]: c = c[ix]
                        c[a[-1]] = V(a[2]); vals[a[1]] = obj
                    elif op == 79:
                        v1, v2 = V(a[2]), V(a[3]); cat = list(v1) + list(v2)
                        vals[a[1]] = [cat[ix] if ix != 0xFFFFFFFF else self._null(types[a[0]][1]) for ix in a[4:]]
                    elif op in (83, 400): vals[a[1]] = V(a[2])
                    elif op == 124:   # Bitcast
                        vals[a[1]] = self._bitcast(a[0], V(a[2]))
                    elif op == 127: vals[a[1]] = vmap(lambda x: F32(-x), V(a[2]))
                    elif op in (129, 131, 133, 136):
                        f = {129: lambda x, y: F32(x + y), 131: lambda x, y: F32(x - y), 133: lambda x, y: F32(x * y), 136: lambda x, y: F32(x / y)}[op]
                        vals[a[1]] = vmap(f, V(a[2]), V(a[3]))
                    elif op in (128, 130, 132, 134, 137, 194, 196, 197, 199, 198):
                        w = width(a[0]); mask = (1 << w) - 1
                        f = {128: lambda x, y: (x + y) & mask, 130: lambda x, y: (x - y) & mask, 132: lambda x, y: (x * y) & mask,
                             134: lambda x, y: (x // y) if y else 0, 137: lambda x, y: (x % y) if y else 0,
                             194: lambda x, y: (x >> y) if y < w else 0, 196: lambda x, y: ((x << y) & mask) if y < w else 0,
                             197: lambda x, y: x | y, 199: lambda x, y: x & y, 198: lambda x, y: x ^ y}[op]
                        vals[a[1]] = vmap(f, V(a[2]), V(a[3]))
                    elif op == 126:
                        w = width(a[0]); vals[a[1]] = vmap(lambda x: (-x) & ((1 << w) - 1), V(a[2]))
                    elif op == 142: vals[a[1]] = [F32(x * V(a[3])) for x in V(a[2])]
                    elif op == 145: vals[a[1]] = mat_vec(V(a[2]), V(a[3]))
                    elif op == 146:
                        A, B = V(a[2]), V(a[3]); vals[a[1]] = [mat_vec(A, col) for col in B]
                    elif op == 148: vals[a[1]] = dot(V(a[2]), V(a[3]))
                    elif op in (164, 166, 167):
                        f = {164: lambda x, y: x == y, 166: lambda x, y: x or y, 167: lambda x, y: x and y}[op]
                        vals[a[1]] = vmap(f, V(a[2]), V(a[3]))
                    elif op == 168: vals[a[1]] = vmap(lambda x: not x, V(a[2]))
                    elif op == 169:
                        c, x, y = V(a[2]), V(a[3]), V(a[4])
                        vals[a[1]] = [xi if ci else yi for ci, xi, yi in zip(c, x, y)] if isinstance(c, list) else (x if c else y)
                    elif op in (170, 171, 172, 174, 176, 178):
                        f = {170: lambda x, y: x == y, 171: lambda x, y: x != y, 172: lambda x, y: x > y, 174: lambda x, y: x >= y,
                             176: lambda x, y: x < y, 178: lambda x, y: x <= y}[op]
                        vals[a[1]] = vmap(f, V(a[2]), V(a[3]))
                    elif op in (173, 175, 177, 179):
                        w = width(m.defs[a[2]].rtype) if a[2] in m.defs and m.defs[a[2]].rtype else 32
                        f = {173: lambda x, y: signed(x, w) > signed(y, w), 175: lambda x, y: signed(x, w) >= signed(y, w),
                             177: lambda x, y: signed(x, w) < signed(y, w), 179: lambda x, y: signed(x, w) <= signed(y, w)}[op]
                        vals[a[1]] = vmap(f, V(a[2]), V(a[3]))
                    elif op in (180, 182, 184, 186, 188, 190):   # ordered comparisons: false on NaN
                        f = {180: lambda x, y: bool(x == y), 182: lambda x, y: bool(x != y) and not (np.isnan(x) or np.isnan(y)),
                             184: lambda x, y: bool(x < y), 186: lambda x, y: bool(x > y), 188: lambda x, y: bool(x <= y), 190: lambda x, y: bool(x >= y)}[op]
                        vals[a[1]] = vmap(f, V(a[2]), V(a[3]))
                    elif op in (185, 187, 189, 191):   # unordered: true on NaN
                        g = {185: lambda x, y: x < y, 187: lambda x, y: x > y, 189: lambda x, y: x <= y, 191: lambda x, y: x >= y}[op]
                        vals[a[1]] = vmap(lambda x, y: bool(np.isnan(x) or np.isnan(y) or g(x, y)), V(a[2]), V(a[3]))
                    elif op == 109:   # ConvertFToU: round toward zero, out of range pinned like DESIGN.md §3
                        w = width(a[0])
                        def f2u(x, w=w):
                            x = float(x)
                            if math.isnan(x) or x <= 0.0: return 0
                            if math.isinf(x) or x >= float(1 << w): return (1 << w) - 1
                            return int(x)
                        vals[a[1]] = vmap(f2u, V(a[2]))
                    elif op == 110:
                        w = width(a[0])
                        def f2s(x, w=w):
                            x = float(x)
                            if math.isnan(x): return 0
                            if math.isinf(x): x = math.copysign(float(1 << w), x)
                            v = max(min(int(x), (1 << (w - 1)) - 1), -(1 << (w - 1)))
                            return v & ((1 << w) - 1)
                        vals[a[1]] = vmap(f2s, V(a[2]))
                    elif op == 111:
                        w = width(m.defs[a[2]].rtype) if a[2] in m.defs and m.defs[a[2]].rtype else 32
                        vals[a[1]] = vmap(lambda x: F32(signed(x, w)), V(a[2]))
                    elif op == 112: vals[a[1]] = vmap(lambda x: F32(x), V(a[2]))
                    elif op == 113:
                        w = width(a[0]); vals[a[1]] = vmap(lambda x: x & ((1 << w) - 1), V(a[2]))
                    elif op == 114:
                        w = width(a[0]); ws = width(m.defs[a[2]].rtype)
                        vals[a[1]] = vmap(lambda x: signed(x, ws) & ((1 << w) - 1), V(a[2]))
                    elif op == 12: vals[a[1]] = self._ext(a, V, vmap, dot)
                    elif op in (234, 239, 241, 237, 240, 242, 235, 229, 230):
                        yield ("sync", inst.index, None)      # scheduling point: see _schedule
                        p = vals[a[2]] if a[2] in vals else var_pointer(a[2])
                        old = load(p)
                        v = V(a[5]) if len(a) > 5 else None
                        new = {234: lambda: (old + v) & 0xFFFFFFFF, 235: lambda: (old - v) & 0xFFFFFFFF, 239: lambda: max(old, v), 237: lambda: min(old, v),
                               241: lambda: old | v, 240: lambda: old & v, 242: lambda: old ^ v, 229: lambda: v, 230: lambda: old}[op]()
                        store(p, new)
                        vals[a[1]] = old
                    elif op == 339:
                        vals[a[1]] = yield ("ballot", inst.index, bool(V(a[3])))
                    elif op == 333:
                        vals[a[1]] = yield ("elect", inst.index, None)
                    elif op == 86: vals[a[1]] = (V(a[2]), V(a[3]))
                    elif op == 100: vals[a[1]] = V(a[2])[0]
                    elif op == 88:
                        img, smp = V(a[2]); coord = V(a[3])
                        assert a[4] & 2, "explicit Lod expected"
                        vals[a[1]] = self._sample(img, smp, coord, V(a[5]))
                    elif op == 87:     # ImageSampleImplicitLod in a compute shader: lod 0
                        img, smp = V(a[2]); vals[a[1]] = self._sample(img, smp, V(a[3]), F32(0))
                    elif op == 95:
                        img = V(a[2]); c = V(a[3]); lod = V(a[5]) if len(a) > 5 and (a[4] & 2) else 0
                        if isinstance(img, tuple): img = img[0]
                        lv = img.levels[lod]
                        x = min(max(signed(c[0], 32), 0), lv.shape[1] - 1); y = min(max(signed(c[1], 32), 0), lv.shape[0] - 1)
                        t = lv[y, x]
                        vals[a[1]] = [F32(t), F32(0), F32(0), F32(1)] if lv.ndim == 2 else list(t) + [0] * (4 - len(t))
                    elif op == 99:
                        img = V(a[0]); c = V(a[1]); t = V(a[2]); lv = img.levels[0]
                        idx = tuple(signed(x, 32) for x in reversed(c)) if isinstance(c, list) else (c,)
                        if lv.ndim == len(idx): lv[idx] = t[0]
                        else: lv[idx] = t[:lv.shape[-1]]
                    elif op == 103:
                        img = V(a[2]); lod = V(a[3])
                        if isinstance(img, tuple): img = img[0]
                        lv = img.levels[lod]; vals[a[1]] = [lv.shape[1], lv.shape[0]]
                    elif op == 249: nxt = a[0]; break
                    elif op == 250: nxt = a[1] if V(a[0]) else a[2]; break
                    elif op == 251:
                        sel = V(a[0]); nxt = a[1]
                        for j in range(2, len(a), 2):
                            if a[j] == sel: nxt = a[j + 1]; break
                        break
                    elif op == 253: return
                    elif op == 255: raise RuntimeError("OpUnreachable executed")
                    elif op == 224:
                        yield ("barrier", inst.index, None)
                    elif op == 5294:        # EmitMeshTasksEXT: ends the invocation; the group's task count + payload are recorded
                        builtins["emit"]((V(a[0]), V(a[1]), V(a[2])), self._copy(load(a[3])) if len(a) > 3 else None)
                        return
                    elif op == 225: pass    # MemoryBarrier
                    else:
                        raise NotImplementedError("opcode %d (line %d)" % (op, inst.line))
                prev, label = label, nxt

    def _construct(self, t, parts):
        ty = self.types[t]
        if ty[0] == "vec":
            out = []
            for p in parts:
                if isinstance(p, list): out.extend(p)
                else: out.append(p)
            return out
        return list(parts)

    def _copy(self, v):
        return [self._copy(x) for x in v] if isinstance(v, list) else v

    def _bitcast(self, t, v):
        ty = self.types[t]
        def one(x, to):
            if to[0] == "float": return F32(struct.unpack("<f", struct.pack("<I", x & 0xFFFFFFFF))[0]) if not isinstance(x, np.floating) else x
            if to[0] == "int": return struct.unpack("<I", struct.pack("<f", float(x)))[0] if isinstance(x, np.floating) else x
            raise NotImplementedError(to)
        if ty[0] == "vec": return [one(x, self.types[ty[1]]) for x in v]
        return one(v, ty)

    def _ext(self, a, V, vmap, dot):
        e = a[3]; x = [V(i) for i in a[4:]]
        if e == 31: return vmap(lambda v: F32(np.sqrt(v)), x[0])
        if e == 32: return vmap(lambda v: F32(F32(1) / F32(np.sqrt(v))), x[0])
        if e == 30: return vmap(lambda v: F32(self.log2f(v)), x[0])
        if e == 4: return vmap(lambda v: F32(abs(v)), x[0])
        if e == 8: return vmap(lambda v: F32(np.floor(v)), x[0])
        if e == 9: return vmap(lambda v: F32(np.ceil(v)), x[0])
        if e in (37, 79): return vmap(lambda p, q: q if (q < p or np.isnan(p)) else p, x[0], x[1])      # fminf semantics
        if e in (40, 80): return vmap(lambda p, q: q if (q > p or np.isnan(p)) else p, x[0], x[1])
        if e == 38: return vmap(min, x[0], x[1])
        if e == 41: return vmap(max, x[0], x[1])
        if e in (43, 81):
            mx = vmap(lambda p, q: q if (q > p or np.isnan(p)) else p, x[0], x[1])
            return vmap(lambda p, q: q if (q < p or np.isnan(p)) else p, mx, x[2])
        if e == 44: return vmap(lambda p, lo, hi: min(max(p, lo), hi), x[0], x[1], x[2])
        if e == 50: return vmap(fma32, x[0], x[1], x[2])
        if e == 66: return F32(np.sqrt(dot(x[0], x[0]))) if isinstance(x[0], list) else F32(abs(x[0]))
        if e == 67:
            d = [F32(p - q) for p, q in zip(x[0], x[1])]
            return F32(np.sqrt(dot(d, d)))
        raise NotImplementedError("GLSL.std.450 %d" % e)

    def _sample(self, img, smp, coord, lod):
        n = len(img.levels)
        lod = float(lod)
        if math.isnan(lod): lvl = 0
        else:
            lod = min(max(lod, 0.0), float(n - 1))
            lvl = int(math.ceil(lod + 0.5)) - 1
            lvl = min(max(lvl, 0), n - 1)
        lv = img.levels[lvl]
        h, w = lv.shape

        def fp(u, size):
            fx = F32(F32(u * F32(size)) - F32(0.5))
            f = np.floor(fx)
            if not (f >= 0): a0 = -1
            elif f >= size: a0 = size
            else: a0 = int(f)
            return min(max(a0, 0), size - 1), min(max(a0 + 1, 0), size - 1)
        x0, x1 = fp(coord[0], w); y0, y1 = fp(coord[1], h)
        if smp.reduce_min:
            t = min(lv[y0, x0], lv[y0, x1], lv[y1, x0], lv[y1, x1])
        else:
            t = lv[y0, x0]
        return [F32(t), F32(0), F32(0), F32(1)]
