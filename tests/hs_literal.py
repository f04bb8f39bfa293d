"""hs_literal.py -- TEST INFRASTRUCTURE: a second, independent restatement of the reference's hot path, written from
the Haskell sources (NOT from oracle/oracle.c) in the most literal Python possible: recursion stays recursion, lists
stay lists, `minimumBy` is `foldr1`, every arithmetic operation is one numpy float32 operation in the order the
Haskell expression tree gives it.  Slow by design (a few thousand rays per second); it exists to pin oracle.c, which
would otherwise only be checked against itself (the reference ships no tests and cannot be built here: no GHC).

Follows, function by function:
  V3.hs:5-37                      V3, Num instance, *^, cross, dot, norm, normalize
  Geometry.hs:79-83,104-107       normal, vertices, rotVert
  Geometry.hs:117-142             mollerTrumbore
  Geometry.hs:155-197             getBounds, intersectsBB, averagePoints, dim, longestAxis, boundingBox, projectToAxis
  BIH.hs:62-96                    makeBIH, bih, split
  BIH.hs:101-141                  intersectBIH, intersectBIH'
  Lib.hs:79-89,93-104,107-114     renderPixel, rgbFloatToPixelRGB, makeRay
  Lib.hs:127-137,155-198          raytrace, bounceRay, scatterRay, reflectRay, randomR, randomVector
base-4.9.1 (GHC 8.0.2, lts-9.8) semantics used:
  Ord Float class defaults   max x y = if x <= y then y else x ; min x y = if x <= y then x else y
  compare on Float           x < y -> LT ; x == y -> EQ ; otherwise GT
  Data.Foldable.minimumBy    foldr1 (\\x y -> case cmp x y of GT -> y ; _ -> x)
  Data.Foldable.maximumBy    foldr1 (\\x y -> case cmp x y of GT -> x ; _ -> y)
  GHC.List maximum/minimum   foldl1 max / foldl1 min ; sum = foldl (+) 0 ; genericLength = foldr (\\_ n -> 1 + n) 0
The generator is not tf-random (un-vendored, replaced by Philox4x32-10 in this project): `TFGen`
below is a counter into the Philox stream with the same interface (mkTFGen seed, next), so the reuse pattern of draws
falls out of the transliteration instead of being coded.  cos/sin/acos/atan are libm's float functions (what GHC calls).
"""
import ctypes
import ctypes.util
import sys

import numpy as np

sys.setrecursionlimit(100000)
f32 = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m"))
for _n in ("cosf", "sinf", "acosf", "atanf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
_err = np.seterr(all="ignore")          # 1/0, 0*inf are values here, not errors


def cos(x): return f32(_libm.cosf(float(x)))
def sin(x): return f32(_libm.sinf(float(x)))
def acos(x): return f32(_libm.acosf(float(x)))
def atan(x): return f32(_libm.atanf(float(x)))


PI = f32(np.pi)                          # pi :: Float


# ---- Prelude pieces ---------------------------------------------------------------------------------------------
def hs_max(x, y): return y if x <= y else x
def hs_min(x, y): return x if x <= y else y


def compare(x, y):
    if x < y:
        return "LT"
    if x == y:
        return "EQ"
    return "GT"


def foldr1(f, xs):
    acc = xs[-1]
    for x in reversed(xs[:-1]):
        acc = f(x, acc)
    return acc


def foldl1(f, xs):
    acc = xs[0]
    for x in xs[1:]:
        acc = f(acc, x)
    return acc


def minimumBy(cmp, xs): return foldr1(lambda x, y: y if cmp(x, y) == "GT" else x, xs)
def maximumBy(cmp, xs): return foldr1(lambda x, y: x if cmp(x, y) == "GT" else y, xs)
def comparing(key): return lambda a, b: compare(key(a), key(b))
def maximum(xs): return foldl1(hs_max, xs)
def minimum(xs): return foldl1(hs_min, xs)
def maximumDef(d, xs): return maximum(xs) if xs else d
def minimumDef(d, xs): return minimum(xs) if xs else d


def signum(x):
    if x > 0:
        return f32(1)
    if x < 0:
        return f32(-1)
    return x


def genericLength(xs):
    n = f32(0)
    for _ in xs:
        n = f32(1) + n
    return n


# ---- V3.hs ------------------------------------------------------------------------------------------------------
class V3:
    __slots__ = ("x", "y", "z")

    def __init__(self, x, y, z):
        self.x, self.y, self.z = f32(x), f32(y), f32(z)

    def __add__(s, o): return V3(s.x + o.x, s.y + o.y, s.z + o.z)
    def __mul__(s, o): return V3(s.x * o.x, s.y * o.y, s.z * o.z)
    def __neg__(s): return V3(-s.x, -s.y, -s.z)
    def __sub__(s, o): return s + (-o)                   # class default: x - y = x + negate y


def fromInteger(n): return V3(n, n, n)
def vmap(f, v): return V3(f(v.x), f(v.y), f(v.z))
def scale(r, v): return V3(r * v.x, r * v.y, r * v.z)            # (*^)


def cross(p, q):
    a, b, c, d, e, f = p.x, p.y, p.z, q.x, q.y, q.z
    return V3(b * f - c * e, c * d - a * f, a * e - b * d)


def dot(p, q): return (p.x * q.x) + (p.y * q.y) + (p.z * q.z)
def norm(v): return np.sqrt(dot(v, v))


def normalize(v):
    vnorm = norm(v)
    return V3(v.x / vnorm, v.y / vnorm, v.z / vnorm)


def vsum(vs):                                             # sum = foldl (+) 0
    acc = fromInteger(0)
    for v in vs:
        acc = acc + v
    return acc


# ---- Geometry.hs ------------------------------------------------------------------------------------------------
class Material:
    def __init__(self, reflective, surfColor, emissive, emitColor):
        self.reflective, self.surfColor, self.emissive, self.emitColor = f32(reflective), surfColor, f32(emissive), emitColor


class Triangle:
    __slots__ = ("tFirst", "tSecond", "tThird", "material", "index")

    def __init__(self, a, b, c, material, index):
        self.tFirst, self.tSecond, self.tThird, self.material, self.index = a, b, c, material, index


class Ray:
    __slots__ = ("vertex", "direction")

    def __init__(self, vertex, direction): self.vertex, self.direction = vertex, direction


class Intersection:
    __slots__ = ("intersectPoint", "dist", "surface")

    def __init__(self, p, d, s): self.intersectPoint, self.dist, self.surface = p, d, s


class Bounds:
    __slots__ = ("lo", "hi")

    def __init__(self, lo, hi): self.lo, self.hi = lo, hi


X, Y, Z = "X", "Y", "Z"
def normal(t): return cross(t.tSecond - t.tFirst, t.tThird - t.tFirst)
def vertices(t): return [t.tFirst, t.tSecond, t.tThird]
def projectToAxis(ax, v): return {X: v.x, Y: v.y, Z: v.z}[ax]
def to(a, b): return Ray(a, b - a)


def mollerTrumbore(ray, tri):
    eps = f32(0.0001)
    rayVert, rayDir = ray.vertex, ray.direction
    vertex0, vertex1, vertex2 = tri.tFirst, tri.tSecond, tri.tThird
    edge1 = vertex1 - vertex0
    edge2 = vertex2 - vertex0
    h = cross(rayDir, edge2)
    a = dot(edge1, h)
    if a > -eps and a < eps:
        return None
    f = f32(1) / a
    s = rayVert - vertex0
    u = f * dot(s, h)
    if u < 0 or u > 1:
        return None
    q = cross(s, edge1)
    v = f * dot(rayDir, q)
    if v < 0 or u + v > 1:
        return None
    t = f * dot(edge2, q)
    if t > eps:
        outInter = rayVert + scale(t, rayDir)
        rayDist = norm(outInter - rayVert)
        return Intersection(outInter, rayDist, tri)
    return None


def naiveIntersect(tris, ray):
    intersections = [i for i in (mollerTrumbore(ray, t) for t in tris) if i is not None]
    if not intersections:
        return None
    return minimumBy(comparing(lambda i: i.dist), intersections)


def getBounds(verts):
    xs, ys, zs = [v.x for v in verts], [v.y for v in verts], [v.z for v in verts]
    return Bounds(V3(minimum(xs), minimum(ys), minimum(zs)), V3(maximum(xs), maximum(ys), maximum(zs)))


def intersectsBB(b, ray):
    lx, ly, lz, hx, hy, hz = b.lo.x, b.lo.y, b.lo.z, b.hi.x, b.hi.y, b.hi.z
    vx, vy, vz = ray.vertex.x, ray.vertex.y, ray.vertex.z
    dfx, dfy, dfz = f32(1) / ray.direction.x, f32(1) / ray.direction.y, f32(1) / ray.direction.z
    t1 = (lx - vx) * dfx
    t2 = (hx - vx) * dfx
    t3 = (ly - vy) * dfy
    t4 = (hy - vy) * dfy
    t5 = (lz - vz) * dfz
    t6 = (hz - vz) * dfz
    tmin = hs_max(hs_max(hs_min(t1, t2), hs_min(t3, t4)), hs_min(t5, t6))
    tmax = hs_min(hs_min(hs_max(t1, t2), hs_max(t3, t4)), hs_max(t5, t6))
    return bool(tmax > 0 and tmin < tmax)


def averagePoints(verts):
    n = genericLength(verts)
    return vmap(lambda c: c / n, vsum(verts))


def dim(ax, b): return projectToAxis(ax, b.hi) - projectToAxis(ax, b.lo)


def longestAxis(b):
    return maximumBy(comparing(lambda p: p[1]), list(zip([X, Y, Z], [dim(X, b), dim(Y, b), dim(Z, b)])))[0]


def boundingBox(tris): return getBounds([v for t in tris for v in vertices(t)])


# ---- BIH.hs -----------------------------------------------------------------------------------------------------
class Leaf:
    def __init__(self, geom): self.geom = geom


class Branch:
    def __init__(self, node, l, r): self.node, self.l, self.r = node, l, r       # node = (axis, lmax, rmin)


class BIH:
    def __init__(self, bounds, tree): self.bounds, self.tree = bounds, tree


def makeBIH(tris):
    bbox = boundingBox(tris)
    return BIH(bbox, bih(bbox, list(tris)))


def bih(bbox, geom):
    leafLimit = 15
    if len(geom) < leafLimit:
        return Leaf(geom)
    leftTris, lmax, rightTris, rmin, axis = split(bbox, list(geom))
    if not leftTris:
        return Branch((axis, lmax, rmin), Leaf([]), Leaf(rightTris))
    if not rightTris:
        return Branch((axis, lmax, rmin), Leaf(leftTris), Leaf([]))
    return Branch((axis, lmax, rmin), bih(boundingBox(leftTris), leftTris), bih(boundingBox(rightTris), rightTris))


def split(bbox, geom):
    lo, hi = bbox.lo, bbox.hi
    ax = longestAxis(bbox)
    splitPlane = projectToAxis(ax, averagePoints([averagePoints(vertices(tri)) for tri in geom]))
    def underSplit(tri): return bool(projectToAxis(ax, averagePoints(vertices(tri))) < splitPlane)
    leftTris = [t for t in geom if underSplit(t)]
    rightTris = [t for t in geom if not underSplit(t)]
    leftSide = projectToAxis(ax, lo)
    rightSide = projectToAxis(ax, hi)
    lmax = f32(0.001) + maximumDef(leftSide, [projectToAxis(ax, v) for t in leftTris for v in vertices(t)])
    rmin = f32(-0.001) + minimumDef(rightSide, [projectToAxis(ax, v) for t in rightTris for v in vertices(t)])
    return leftTris, lmax, rightTris, rmin, ax


def height(t): return 1 if isinstance(t, Leaf) else 1 + max(height(t.l), height(t.r))
def numLeaves(t): return 1 if isinstance(t, Leaf) else numLeaves(t.l) + numLeaves(t.r)
def longestLeaf(t): return len(t.geom) if isinstance(t, Leaf) else max(longestLeaf(t.l), longestLeaf(t.r))
def flatten(t): return list(t.geom) if isinstance(t, Leaf) else flatten(t.l) + flatten(t.r)


def intersectBIH(b, ray): return intersectBIH_(b.bounds, b.tree, ray)


def intersectBIH_(bbox, tree, ray):
    if isinstance(tree, Leaf):
        intersections = [i for i in (mollerTrumbore(ray, t) for t in tree.geom) if i is not None]      # V.mapMaybe
        if not intersections:
            return None
        return minimumBy(comparing(lambda i: i.dist), intersections)
    ax, lmax, rmin = tree.node
    l, r = tree.l, tree.r
    if not intersectsBB(bbox, ray):
        return None
    low, high = bbox.lo, bbox.hi
    left = Bounds(low, {X: V3(lmax, high.y, high.z), Y: V3(high.x, lmax, high.z), Z: V3(high.x, high.y, lmax)}[ax])
    right = Bounds({X: V3(rmin, low.y, low.z), Y: V3(low.x, rmin, low.z), Z: V3(low.x, low.y, rmin)}[ax], high)
    leftToRight = bool(projectToAxis(ax, ray.direction) > 0)
    intersectsLeft = intersectsBB(left, ray)
    intersectsRight = intersectsBB(right, ray)

    def isClose(v):
        if leftToRight:
            return bool(projectToAxis(ax, v.intersectPoint) < rmin)
        return bool(projectToAxis(ax, v.intersectPoint) > lmax)
    if intersectsLeft and intersectsRight:
        # [near, far] are lazy in Haskell: far is only forced where the case expression needs it
        near = intersectBIH_(left, l, ray) if leftToRight else intersectBIH_(right, r, ray)
        def far(): return intersectBIH_(right, r, ray) if leftToRight else intersectBIH_(left, l, ray)
        if near is not None:
            if isClose(near):
                return near
            intersections = [i for i in (near, far()) if i is not None]          # catMaybes [near, far]
            return minimumBy(comparing(lambda i: i.dist), intersections)         # minimumByMay on a non-empty list
        return far()
    if intersectsLeft:
        return intersectBIH_(left, l, ray)
    if intersectsRight:
        return intersectBIH_(right, r, ray)
    return None


# ---- the generator interface Lib.hs uses (tf-random's, over this project's Philox stream) ---------------------------
def _philox4x32_10(c, k):
    c0, c1, c2, c3 = c
    k0, k1 = k
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xffffffff, p1 & 0xffffffff, ((p0 >> 32) ^ c3 ^ k1) & 0xffffffff, p0 & 0xffffffff
        k0, k1 = (k0 + 0x9E3779B9) & 0xffffffff, (k1 + 0xBB67AE85) & 0xffffffff
    return c0, c1, c2, c3


class TFGen:
    key = (0, 0)                                           # the render's seed (sqt_render_params.seed)

    def __init__(self, stream, j=0): self.stream, self.j = stream, j

    def next(self):
        w = _philox4x32_10((self.stream & 0xffffffff, (self.stream >> 32) & 0xffffffff, self.j >> 2, 0x52545153), TFGen.key)
        return w[self.j & 3], TFGen(self.stream, self.j + 1)


def mkTFGen(n): return TFGen(n)


# ---- Lib.hs -----------------------------------------------------------------------------------------------------
def randomR(lohi, g):
    lo, hi = f32(lohi[0]), f32(lohi[1])
    n, g2 = g.next()
    p = f32(n) / f32(4294967295)             # fromIntegral n / fromIntegral (maxBound :: Word32), both rounded to Float
    r = hi - lo
    return lo + r * p, g2


def randomVector(gen):
    u, gen2 = randomR((0, 1), gen)
    v, _ = randomR((0, 1), gen2)
    th = f32(2) * PI * u
    ph = acos(f32(2) * v - f32(1))
    return V3(cos(th) * sin(ph), sin(th) * sin(ph), cos(ph))


def reflectRay(ray, inter):
    dn = normalize(normal(inter.surface))
    di = ray.direction
    newDir = di - scale(f32(2) * dot(dn, di), dn)
    return Ray(inter.intersectPoint, newDir)


def scatterRay(gen, ray, inter):
    newDir = randomVector(gen)
    old = signum(dot(ray.direction, normal(inter.surface)))
    new = signum(dot(newDir, normal(inter.surface)))
    if old == new:
        return Ray(inter.intersectPoint, -newDir)
    return Ray(inter.intersectPoint, newDir)


def bounceRay(gen, ray, inter):
    ref = inter.surface.material.reflective
    x = randomR((0, 1), gen)[0]
    if ref < x:
        return scatterRay(gen, ray, inter)
    return reflectRay(ray, inter)


black = fromInteger(0)


def raytrace(gen, isect, ray, bounces, max_bounces=2):
    """Lib.hs:127-137; `max_bounces` = the literal 2 of `bounces > 2` (an extension of this project makes it a parameter)."""
    if bounces > max_bounces:
        return black
    inter = isect(ray)
    if inter is None:
        return black
    m = inter.surface.material
    newRay = bounceRay(gen, ray, inter)
    newGen = gen.next()[1]
    nextBounce = m.surfColor * raytrace(newGen, isect, newRay, bounces + 1, max_bounces)
    emitContribution = scale(m.emissive, m.emitColor)
    return nextBounce + emitContribution


def rotVert(vert, matr):
    """fromV vert * matr with Data.Matrix multStd: element (1,j) = sum [ a!(1,k) * b!(k,j) | k <- [1..3] ], sum = foldl (+) 0"""
    row = [vert.x, vert.y, vert.z]
    out = []
    for j in range(3):
        acc = f32(0)
        for k in range(3):
            acc = acc + row[k] * f32(matr[k][j])
        out.append(acc)
    return V3(*out)


def makeRay(dims, ix, cam_pos, cam_rot):
    w, h = dims
    y, x = ix
    ww, hh = f32(w), f32(h)
    xoffs = (f32(x) - (ww / f32(2))) / ww
    yoffs = ((hh / f32(2)) - f32(y)) / hh
    return Ray(cam_pos, rotVert(V3(1, xoffs, yoffs), cam_rot))


def renderPixelSum(isect, cam_pos, cam_rot, sampleCount, dims, ix, max_bounces=2, seed_stride=None):
    """Lib.hs:79-88 up to `sum outcomes` (the per-pixel radiance sum the C ABI exposes as accum_out)."""
    w, h = dims
    y, x = ix
    ray = makeRay(dims, ix, cam_pos, cam_rot)
    rix = sampleCount * (x + y * (w if seed_stride is None else seed_stride))
    rngs = [mkTFGen(rix + k) for k in range(sampleCount)]
    outcomes = [raytrace(r, isect, ray, 0, max_bounces) for r in rngs]
    return vsum(outcomes)


def _floor_word8(x):
    """floor :: Float -> Word8 through properFraction/decodeFloat (wraps modulo 256; NaN and infinities give 0)."""
    x = float(x)
    if x != x or x in (float("inf"), float("-inf")):
        return 0
    n = int(x)                                # truncation towards zero (quotRem)
    r = x - n
    return ((n - 1) if r < 0 else n) % 256


def rgbFloatToPixelRGB(c):
    r, g, b = c.x, c.y, c.z
    maxComponent = hs_max(hs_max(r, g), b)
    minComponent = hs_min(hs_min(r, g), b)
    lightness = f32(0.5) * (maxComponent + minComponent)
    intensity = atan(lightness) / (PI / f32(2))
    s = scale(intensity / maxComponent, c)
    return tuple(min(255, _floor_word8(v * f32(255))) for v in (s.x, s.y, s.z))


# ---- glue for the tests -----------------------------------------------------------------------------------------
def triangles_from_arrays(v9, mat_idx, mats8):
    mats = [Material(m[0], V3(m[1], m[2], m[3]), m[4], V3(m[5], m[6], m[7])) for m in np.asarray(mats8, np.float32)]
    v9 = np.asarray(v9, np.float32).reshape(-1, 9)
    return [Triangle(V3(*v[0:3]), V3(*v[3:6]), V3(*v[6:9]), mats[int(m)], i) for i, (v, m) in enumerate(zip(v9, mat_idx))]


def serialize_tree(b):
    """(bounds, preorder list of ('B', axis, lmax bits, rmin bits) / ('L', [triangle indices])) for comparison with the oracle's tree"""
    out = []

    def walk(t):
        if isinstance(t, Leaf):
            out.append(("L", [tri.index for tri in t.geom]))
        else:
            ax, lmax, rmin = t.node
            out.append(("B", "XYZ".index(ax), int(np.float32(lmax).view(np.uint32)), int(np.float32(rmin).view(np.uint32))))
            walk(t.l); walk(t.r)
    walk(b.tree)
    bounds = [float(v) for v in (b.bounds.lo.x, b.bounds.lo.y, b.bounds.lo.z, b.bounds.hi.x, b.bounds.hi.y, b.bounds.hi.z)]
    return bounds, out
