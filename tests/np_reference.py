"""Independent numpy float32 restatement of the path (a second implementation of
the same specification, written array-at-a-time).  Used to cross-check the C
oracle; numpy float32 arithmetic is IEEE-754 correctly rounded per operation and
never contracts a*b+c.  fmaf (single rounding) is emulated in float64: the
product of two float32 is exact in float64 and the following add is exact or
correctly rounded well inside float64 for the magnitudes the scenes use.
"""
import numpy as np

f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def _mat_mul(a, b):
    """glm mat4*mat4 on [n,16] column-major arrays, columns accumulated x->w."""
    out = np.empty_like(a)
    for j in range(4):
        for r in range(4):
            s = a[:, r] * b[:, j * 4] + a[:, 4 + r] * b[:, j * 4 + 1]
            s = s + a[:, 8 + r] * b[:, j * 4 + 2]
            out[:, j * 4 + r] = s + a[:, 12 + r] * b[:, j * 4 + 3]
    return out


def trs(pos, rot, scale):
    n = len(pos)
    one, zero, two = f32(1), f32(0), f32(2)
    x, y, z, w = (rot[:, i] for i in range(4))
    np.seterr(all="ignore")
    qxx, qyy, qzz = x * x, y * y, z * z
    qxz, qxy, qyz = x * z, x * y, y * z
    qwx, qwy, qwz = w * x, w * y, w * z
    R = np.zeros((n, 16), f32)
    R[:, 0] = one - two * (qyy + qzz); R[:, 1] = two * (qxy + qwz); R[:, 2] = two * (qxz - qwy)
    R[:, 4] = two * (qxy - qwz); R[:, 5] = one - two * (qxx + qzz); R[:, 6] = two * (qyz + qwx)
    R[:, 8] = two * (qxz + qwy); R[:, 9] = two * (qyz - qwx); R[:, 10] = one - two * (qxx + qyy)
    R[:, 15] = one
    I = np.zeros((n, 16), f32)
    I[:, [0, 5, 10, 15]] = one
    T = I.copy()
    with np.errstate(all="ignore"):
        for r in range(4):
            T[:, 12 + r] = ((I[:, r] * pos[:, 0] + I[:, 4 + r] * pos[:, 1]) + I[:, 8 + r] * pos[:, 2]) + I[:, 12 + r]
        S = I.copy()
        for c in range(3):
            for r in range(4):
                S[:, c * 4 + r] = I[:, c * 4 + r] * scale[:, c]
        return _mat_mul(_mat_mul(T, R), S)


def hierarchy(pos, rot, scale, parent):
    local = trs(pos, rot, scale)
    n = len(pos)
    depth = np.full(n, -1, np.int64)
    for i in range(n):
        chain, cur = [], i
        while depth[cur] < 0:
            chain.append(cur)
            if parent[cur] == 0xFFFFFFFF:
                break
            cur = int(parent[cur])
        d = depth[cur] + 1 if depth[cur] >= 0 and (not chain or chain[-1] != cur) else 0
        for s in reversed(chain):
            depth[s] = d
            d += 1
    world = local.copy()
    for lvl in range(1, int(depth.max()) + 1):
        idx = np.nonzero(depth == lvl)[0]
        with np.errstate(all="ignore"):
            world[idx] = _mat_mul(world[parent[idx].astype(np.int64)], local[idx])
    return world


def view_box(clip):
    """World-space box of the view volume; same formula order as the specification (doubles)."""
    m = [float(x) for x in clip]
    a = [0.0] * 16
    a[0]  =  m[5]*m[10]*m[15] - m[5]*m[11]*m[14] - m[9]*m[6]*m[15] + m[9]*m[7]*m[14] + m[13]*m[6]*m[11] - m[13]*m[7]*m[10]
    a[4]  = -m[4]*m[10]*m[15] + m[4]*m[11]*m[14] + m[8]*m[6]*m[15] - m[8]*m[7]*m[14] - m[12]*m[6]*m[11] + m[12]*m[7]*m[10]
    a[8]  =  m[4]*m[9]*m[15]  - m[4]*m[11]*m[13] - m[8]*m[5]*m[15] + m[8]*m[7]*m[13] + m[12]*m[5]*m[11] - m[12]*m[7]*m[9]
    a[12] = -m[4]*m[9]*m[14]  + m[4]*m[10]*m[13] + m[8]*m[5]*m[14] - m[8]*m[6]*m[13] - m[12]*m[5]*m[10] + m[12]*m[6]*m[9]
    a[1]  = -m[1]*m[10]*m[15] + m[1]*m[11]*m[14] + m[9]*m[2]*m[15] - m[9]*m[3]*m[14] - m[13]*m[2]*m[11] + m[13]*m[3]*m[10]
    a[5]  =  m[0]*m[10]*m[15] - m[0]*m[11]*m[14] - m[8]*m[2]*m[15] + m[8]*m[3]*m[14] + m[12]*m[2]*m[11] - m[12]*m[3]*m[10]
    a[9]  = -m[0]*m[9]*m[15]  + m[0]*m[11]*m[13] + m[8]*m[1]*m[15] - m[8]*m[3]*m[13] - m[12]*m[1]*m[11] + m[12]*m[3]*m[9]
    a[13] =  m[0]*m[9]*m[14]  - m[0]*m[10]*m[13] - m[8]*m[1]*m[14] + m[8]*m[2]*m[13] + m[12]*m[1]*m[10] - m[12]*m[2]*m[9]
    a[2]  =  m[1]*m[6]*m[15]  - m[1]*m[7]*m[14]  - m[5]*m[2]*m[15] + m[5]*m[3]*m[14] + m[13]*m[2]*m[7]  - m[13]*m[3]*m[6]
    a[6]  = -m[0]*m[6]*m[15]  + m[0]*m[7]*m[14]  + m[4]*m[2]*m[15] - m[4]*m[3]*m[14] - m[12]*m[2]*m[7]  + m[12]*m[3]*m[6]
    a[10] =  m[0]*m[5]*m[15]  - m[0]*m[7]*m[13]  - m[4]*m[1]*m[15] + m[4]*m[3]*m[13] + m[12]*m[1]*m[7]  - m[12]*m[3]*m[5]
    a[14] = -m[0]*m[5]*m[14]  + m[0]*m[6]*m[13]  + m[4]*m[1]*m[14] - m[4]*m[2]*m[13] - m[12]*m[1]*m[6]  + m[12]*m[2]*m[5]
    a[3]  = -m[1]*m[6]*m[11]  + m[1]*m[7]*m[10]  + m[5]*m[2]*m[11] - m[5]*m[3]*m[10] - m[9]*m[2]*m[7]   + m[9]*m[3]*m[6]
    a[7]  =  m[0]*m[6]*m[11]  - m[0]*m[7]*m[10]  - m[4]*m[2]*m[11] + m[4]*m[3]*m[10] + m[8]*m[2]*m[7]   - m[8]*m[3]*m[6]
    a[11] = -m[0]*m[5]*m[11]  + m[0]*m[7]*m[9]   + m[4]*m[1]*m[11] - m[4]*m[3]*m[9]  - m[8]*m[1]*m[7]   + m[8]*m[3]*m[5]
    a[15] =  m[0]*m[5]*m[10]  - m[0]*m[6]*m[9]   - m[4]*m[1]*m[10] + m[4]*m[2]*m[9]  + m[8]*m[1]*m[6]   - m[8]*m[2]*m[5]
    det = m[0]*a[0] + m[1]*a[4] + m[2]*a[8] + m[3]*a[12]
    inf = float("inf")
    ok = det != 0.0 and np.isfinite(det)
    lo, hi = [inf] * 3, [-inf] * 3
    for c in range(8):
        if not ok:
            break
        x, y, z = (1.0 if c & 1 else -1.0), (1.0 if c & 2 else -1.0), (1.0 if c & 4 else -1.0)
        w = (a[3]*x + a[7]*y + a[11]*z + a[15]) / det
        if not (w > 0.0) or not np.isfinite(w):
            ok = False
            break
        p = [(a[0]*x + a[4]*y + a[8]*z + a[12]) / det / w, (a[1]*x + a[5]*y + a[9]*z + a[13]) / det / w,
             (a[2]*x + a[6]*y + a[10]*z + a[14]) / det / w]
        for i in range(3):
            if not np.isfinite(p[i]):
                ok = False
            lo[i], hi[i] = min(lo[i], p[i]), max(hi[i], p[i])
    if not ok:
        return np.full(3, -np.inf, f32), np.full(3, np.inf, f32)
    out_lo, out_hi = np.empty(3, f32), np.empty(3, f32)
    for i in range(3):
        d = (hi[i] - lo[i]) * 1e-4 + 1e-6
        out_lo[i] = np.nextafter(f32(lo[i] - d), f32(-np.inf))
        out_hi[i] = np.nextafter(f32(hi[i] + d), f32(np.inf))
    return out_lo, out_hi


def extract_view(clip):
    row = [np.array([clip[k * 4 + i] for k in range(4)], f32) for i in range(4)]
    planes = np.stack([row[3] + row[0], row[3] - row[0], row[3] + row[1], row[3] - row[1],
                       row[3] + row[2], row[3] - row[2]]).astype(f32)
    n = planes[:, :3]
    length = np.sqrt((n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1]) + n[:, 2] * n[:, 2]).astype(f32)
    depth = (planes[4] / length[4]).astype(f32)
    rng = f32(depth[3] + planes[5][3] / length[5])
    with np.errstate(all="ignore"):
        scale = f32(16384.0) / rng
    if not (rng > 0 and np.isfinite(scale)):
        scale = f32(0)
    return planes, np.abs(n).astype(f32), length, depth, f32(scale)


def cull_keys(world, mesh, shader, flags, bounds, clip, masks, n_worlds=1, epw=0):
    """Sorted keys and counts [n_worlds*F] (numpy restatement of the spec in SURVEY.md Appendix B)."""
    n, F = len(world), len(masks)
    bounds = np.asarray(bounds, f32).reshape(-1, 7)
    clip = np.asarray(clip, f32).reshape(-1, F, 16)
    E = n if n_worlds == 1 else epw
    wbits = int(np.ceil(np.log2(n_worlds))) if n_worlds > 1 else 0
    valid = mesh < len(bounds)
    b = bounds[np.minimum(mesh, len(bounds) - 1)]
    m = world
    with np.errstate(all="ignore"):
        c = [fma(m[:, 0 + i], b[:, 0], fma(m[:, 4 + i], b[:, 1], fma(m[:, 8 + i], b[:, 2], m[:, 12 + i]))) for i in range(3)]
        e = [fma(np.abs(m[:, 0 + i]), b[:, 3], fma(np.abs(m[:, 4 + i]), b[:, 4], np.abs(m[:, 8 + i]) * b[:, 5])) for i in range(3)]
        ln = [fma(m[:, 4 * j], m[:, 4 * j], fma(m[:, 4 * j + 1], m[:, 4 * j + 1], m[:, 4 * j + 2] * m[:, 4 * j + 2])) for j in range(3)]
        r = b[:, 6] * np.sqrt(np.fmax(np.fmax(ln[0], ln[1]), ln[2])).astype(f32)
    keys, counts = [], np.zeros(n_worlds * F, np.uint32)
    idx_all = np.arange(n, dtype=np.uint64)
    world_of = (idx_all // np.uint64(E)).astype(np.int64) if n_worlds > 1 else np.zeros(n, np.int64)
    for w in range(n_worlds):
        sel_w = world_of == w
        for v in range(F):
            planes, aplanes, length, depth, scale = extract_view(clip[w, v])
            box_lo, box_hi = view_box(clip[w, v])
            part = ((flags & 1) != 0) & ((((flags >> 1) & 15) & int(masks[v])) != 0) & valid & sel_w
            inside = part.copy()
            with np.errstate(all="ignore"):
                for i in range(3):
                    t = e[i] + r
                    inside &= ~((c[i] + t) < box_lo[i])
                    inside &= ~((c[i] - t) > box_hi[i])
                for p in range(6):
                    bc = lambda s: np.full(n, s, f32)
                    dist = fma(bc(planes[p, 0]), c[0], fma(bc(planes[p, 1]), c[1], fma(bc(planes[p, 2]), c[2], bc(planes[p, 3]))))
                    rad = fma(bc(aplanes[p, 0]), e[0], fma(bc(aplanes[p, 1]), e[1], fma(bc(aplanes[p, 2]), e[2], r * length[p])))
                    inside &= ~((dist + rad) < 0)
                zd = fma(bc(depth[0]), c[0], fma(bc(depth[1]), c[1], fma(bc(depth[2]), c[2], bc(depth[3]))))
                t = zd * scale
            d14 = np.where(~(t >= 0), 0, np.where(t >= 16383, 16383, np.nan_to_num(t, nan=0.0, posinf=0, neginf=0).astype(np.int64))).astype(np.uint64)
            sel = np.nonzero(inside)[0]
            local = (sel.astype(np.uint64) - np.uint64(w * E))
            k = np.full(len(sel), w, np.uint64)
            k = (k << np.uint64(5)) | np.uint64(v)
            k = (k << np.uint64(8)) | (shader[sel].astype(np.uint64) & np.uint64(255))
            k = (k << np.uint64(11)) | (mesh[sel].astype(np.uint64) & np.uint64(2047))
            k = (k << np.uint64(14)) | d14[sel]
            k = (k << np.uint64(26 - wbits)) | local
            keys.append(k)
            counts[w * F + v] = len(sel)
    keys = np.sort(np.concatenate(keys)) if keys else np.empty(0, np.uint64)
    return keys, counts


def draw_runs(keys, world_bits=0):
    """f2: the sorted key list cut into runs of one class (world, view, shader, mesh = everything above depth and slot).
    Returns (first position, length, class) per run, in list order."""
    keys = np.asarray(keys, np.uint64)
    if len(keys) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.uint64)
    cls = keys >> np.uint64(40 - world_bits)
    heads = np.flatnonzero(np.concatenate([[True], cls[1:] != cls[:-1]]))
    return heads.astype(np.int64), np.diff(np.concatenate([heads, [len(keys)]])).astype(np.int64), cls[heads]


def indirect_commands(keys, mesh_info, world_bits=0):
    """f2 on the device: one GL DrawElementsIndirectCommand per run -> [n, 5] = count, instanceCount, firstIndex,
    baseVertex, baseInstance.  mesh_info: [n_meshes, 3] = index count, first index, base vertex; a run whose mesh id has
    no entry draws nothing (count 0)."""
    first, length, cls = draw_runs(keys, world_bits)
    mesh_info = np.asarray(mesh_info, np.int64).reshape(-1, 3)
    mesh = (cls & np.uint64(2047)).astype(np.int64)
    known = mesh < len(mesh_info)
    out = np.zeros((len(first), 5), np.int64)
    out[:, 1], out[:, 4] = length, first
    out[known, 0], out[known, 2], out[known, 3] = (mesh_info[mesh[known], c] for c in range(3))
    return out
