/*
 * oracle.c -- CPU restatement of the squigly-trace render hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker or as the timed CPU baseline -- never behind the C-ABI
 * of the CUDA backend.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer values
 * (test/Spec.hs:1-2) and cannot be compiled in this image (no ghc/stack/cabal).  This file is a
 * literal, function-by-function restatement of the Haskell sources, every function citing the
 * lines it follows.  What checks it: the reference's own differential pair (naiveIntersect vs
 * intersectBIH); a low-frequency comparison with render/example.png; and, since round 2, a second
 * restatement written independently from the Haskell (tests/hs_literal.py: literal recursion and
 * lists in Python float32) that must agree with this file bit for bit on the BIH, on 10 k rays
 * incl. the adversarial set, and on the raytrace fold (tests/test_hs_literal.py).  The pin by the
 * reference itself is prepared (tests/golden/ghc/Dump.hs + test_ghc_fixture) but needs GHC.
 *
 * Arithmetic: GHC 8.0.2 Float = IEEE binary32 on SSE scalar, no fusion.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math   (see oracle/Makefile)
 *
 * Deliberate, documented departures (none touches intersection arithmetic):
 *   - RNG: tf-random (Threefish, un-vendored, lts-9.8) is replaced by Philox4x32-10 keyed by
 *     the reference's own generator seed `spp*(x+y*w)+k` (Lib.hs:85-86).  The draw *structure*
 *     of Lib.hs:155-198 (which draw feeds which decision) is kept exactly.
 *   - trig mode 0 calls libm cosf/sinf/acosf/atanf like GHC does; trig mode 1 uses the
 *     polynomial routines the CUDA backend specifies in DESIGN.md ("sqt trig"), restated here
 *     independently so the GPU image can be compared bit-for-bit.
 *   - max_depth is a parameter (reference hard-codes 3, Lib.hs:129).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

/* ------------------------------------------------------------------ V3.hs */
typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
/* V3.hs:8 */
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
/* V3.hs:9 (component-wise product) */
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
/* a - b = a + negate b in the Num default; IEEE a + (-b) == a - b bit for bit */
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
/* V3.hs:18-19 */
static inline v3 vscale(float r, v3 a) { return V(r * a.x, r * a.y, r * a.z); }
/* V3.hs:21-22 */
static inline v3 vcross(v3 p, v3 q)
{
    float a = p.x, b = p.y, c = p.z, d = q.x, e = q.y, f = q.z;
    return V(b * f - c * e, c * d - a * f, a * e - b * d);
}
/* V3.hs:25-26 */
static inline float vdot(v3 p, v3 q) { return (p.x * q.x) + (p.y * q.y) + (p.z * q.z); }
/* V3.hs:31-32 */
static inline float vnorm(v3 a) { return sqrtf(vdot(a, a)); }
/* V3.hs:34-37 */
static inline v3 vnormalize(v3 a) { float n = vnorm(a); return V(a.x / n, a.y / n, a.z / n); }

/* Haskell Ord Float class defaults: max x y = if x <= y then y else x */
static inline float hs_max(float x, float y) { return (x <= y) ? y : x; }
static inline float hs_min(float x, float y) { return (x <= y) ? x : y; }
/* signum for Float: x>0 -> 1 ; x<0 -> -1 ; otherwise x (keeps 0/-0/NaN) */
static inline float hs_signum(float x) { return x > 0 ? 1.0f : (x < 0 ? -1.0f : x); }
static inline float axis_of(v3 a, int ax) { return ax == 0 ? a.x : (ax == 1 ? a.y : a.z); }

/* ------------------------------------------------------------ Color.hs:78-83 */
typedef struct { float reflective; v3 surf; float emissive; v3 emit; } material;

/* --------------------------------------------------------- Geometry.hs types */
typedef struct { v3 a, b, c; int mat; } triangle;      /* Geometry.hs:49-54 (material by index) */
typedef struct { v3 o, d; } ray;                        /* Geometry.hs:44-47 */
typedef struct { v3 lo, hi; } bounds;                   /* Geometry.hs:153 */
typedef struct { int hit; v3 point; float dist; int tri; } isect;   /* Geometry.hs:71-75 */

/* BIH.hs:26,37-43 : Tree BIHNode (Vector Triangle) stored in arrays */
typedef struct {
    int leaf;            /* 1 = Leaf, 0 = Branch */
    int axis; float lmax, rmin; int left, right;     /* Branch (BIHN ax lmax rmin) l r */
    int first, count;    /* Leaf: range in leaf_tris[] */
} bnode;

typedef struct { v3 c; float r; int mat; } sphere;     /* EXTENSION: no counterpart in the reference (see include/sqt.h) */

typedef struct {
    triangle *tris; int n_tris;
    material *mats; int n_mats;
    sphere *spheres; int n_spheres;
    /* BIH */
    bounds root; bnode *nodes; int n_nodes, cap_nodes;
    int *leaf_tris; int n_leaf_tris;          /* original triangle indices in `flatten` order (BIH.hs:50-52) */
    int height, longest_leaf, n_leaves;
    char err[256];
} orc_scene;

typedef struct { uint64_t branch_visits, child_box_tests, own_box_tests, tri_tests, rays; } orc_counters;

/* ===================================================================== Obj.hs */
typedef struct { const char *p; } cursor;
static void skip_spaces(cursor *c) { while (*c->p && isspace((unsigned char)*c->p)) c->p++; }  /* parsec `spaces` */
static int lit(cursor *c, const char *s)
{
    size_t n = strlen(s);
    if (strncmp(c->p, s, n) == 0) { c->p += n; return 1; }
    return 0;
}
/* Obj.hs:131-132 : many1 (noneOf whitespace) <* spaces */
static int word(cursor *c, char *out, size_t cap)
{
    size_t n = 0;
    while (*c->p && !isspace((unsigned char)*c->p)) { if (n + 1 < cap) out[n++] = *c->p; c->p++; }
    out[n] = 0; skip_spaces(c);
    return n > 0;
}
/* Obj.hs:115-121 : optional '-', digits, optional '.', digits ; `read` = correctly rounded */
static int fractional(cursor *c, float *out)
{
    char buf[128]; size_t n = 0; const char *q = c->p;
    if (*q == '-') buf[n++] = *q++;
    while (isdigit((unsigned char)*q) && n < 100) buf[n++] = *q++;
    if (*q == '.') buf[n++] = *q++;
    while (isdigit((unsigned char)*q) && n < 120) buf[n++] = *q++;
    buf[n] = 0;
    if (n == 0 || (n == 1 && (buf[0] == '-' || buf[0] == '.'))) return 0;
    *out = strtof(buf, NULL);
    c->p = q;
    return 1;
}
/* Obj.hs:166-171 */
static int vec3(cursor *c, v3 *out)
{
    if (!fractional(c, &out->x)) return 0; skip_spaces(c);
    if (!fractional(c, &out->y)) return 0; skip_spaces(c);
    if (!fractional(c, &out->z)) return 0; skip_spaces(c);
    return 1;
}

static char *slurp(const char *path)
{
    FILE *f = fopen(path, "rb"); if (!f) return NULL;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    char *b = (char *)malloc((size_t)n + 1);
    if (fread(b, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(b); return NULL; }
    b[n] = 0; fclose(f); return b;
}

typedef struct { char name[128]; material m; } named_mat;
typedef struct { int v0, nv; char mtl[128]; int f0, nf; } object;

/* Obj.hs:49-58,73-86,96-161.  data_dir stands for the hard-coded "./data/" of Obj.hs:52. */
orc_scene *orc_load_obj(const char *obj_path, const char *data_dir)
{
    orc_scene *s = (orc_scene *)calloc(1, sizeof *s);
    char *txt = slurp(obj_path);
    if (!txt) { snprintf(s->err, sizeof s->err, "cannot read %s", obj_path); return s; }
    cursor c = { txt };
    char mtllib[128];
    /* Obj.hs:128-129 */
    if (!lit(&c, "mtllib")) { snprintf(s->err, sizeof s->err, "expected mtllib"); free(txt); return s; }
    skip_spaces(&c); word(&c, mtllib, sizeof mtllib);

    size_t cap_v = 1024, nv = 0, cap_f = 1024, nf = 0, cap_o = 16, no = 0;
    v3 *verts = (v3 *)malloc(cap_v * sizeof *verts);
    int *faces = (int *)malloc(cap_f * 3 * sizeof *faces);
    object *objs = (object *)malloc(cap_o * sizeof *objs);
    /* Obj.hs:99-107 : many parseObj */
    while (*c.p == 'o') {
        c.p++; skip_spaces(&c);
        while (*c.p && (isalnum((unsigned char)*c.p) || *c.p == '.' || *c.p == '_')) c.p++;   /* objectName */
        skip_spaces(&c);
        if (no == cap_o) { cap_o *= 2; objs = (object *)realloc(objs, cap_o * sizeof *objs); }
        object *o = &objs[no++];
        o->v0 = (int)nv; o->nv = 0; o->f0 = (int)nf; o->nf = 0;
        /* Obj.hs:109-113 : 'v' then vec3, stored swapYZ */
        while (c.p[0] == 'v') {
            c.p++; skip_spaces(&c);
            v3 p; if (!vec3(&c, &p)) { snprintf(s->err, sizeof s->err, "bad vertex"); goto done; }
            if (nv == cap_v) { cap_v *= 2; verts = (v3 *)realloc(verts, cap_v * sizeof *verts); }
            verts[nv++] = V(p.x, p.z, p.y); o->nv++;
        }
        /* Obj.hs:125-126 */
        if (!lit(&c, "usemtl")) { snprintf(s->err, sizeof s->err, "expected usemtl"); goto done; }
        skip_spaces(&c); word(&c, o->mtl, sizeof o->mtl);
        /* Obj.hs:134-135 : optional "s on" / "s off" */
        if (lit(&c, "s on") || lit(&c, "s off")) skip_spaces(&c);
        /* Obj.hs:137-147 */
        while (c.p[0] == 'f') {
            c.p++; skip_spaces(&c);
            int idx[3];
            for (int k = 0; k < 3; k++) {
                if (!isdigit((unsigned char)*c.p)) { snprintf(s->err, sizeof s->err, "bad face"); goto done; }
                long v = 0; while (isdigit((unsigned char)*c.p)) { v = v * 10 + (*c.p - '0'); c.p++; }
                skip_spaces(&c); idx[k] = (int)v;
            }
            if (nf == cap_f) { cap_f *= 2; faces = (int *)realloc(faces, cap_f * 3 * sizeof *faces); }
            faces[3 * nf] = idx[0]; faces[3 * nf + 1] = idx[1]; faces[3 * nf + 2] = idx[2];
            nf++; o->nf++;
        }
    }
    {
        /* Obj.hs:52 : material file is looked up under ./data/ whatever the obj path was */
        char mpath[1024]; snprintf(mpath, sizeof mpath, "%s/%s", data_dir, mtllib);
        char *mt = slurp(mpath);
        if (!mt) { snprintf(s->err, sizeof s->err, "cannot read %s", mpath); goto done; }
        cursor m = { mt };
        size_t cap_m = 16, nm = 0; named_mat *mats = (named_mat *)malloc(cap_m * sizeof *mats);
        /* Obj.hs:149-161 */
        while (lit(&m, "newmtl ")) {
            if (nm == cap_m) { cap_m *= 2; mats = (named_mat *)realloc(mats, cap_m * sizeof *mats); }
            named_mat *nmx = &mats[nm];
            word(&m, nmx->name, sizeof nmx->name); skip_spaces(&m);
            if (!lit(&m, "reflective ") || !fractional(&m, &nmx->m.reflective)) break;
            skip_spaces(&m); if (!vec3(&m, &nmx->m.surf)) break; skip_spaces(&m);
            if (!lit(&m, "emissive ") || !fractional(&m, &nmx->m.emissive)) break;
            skip_spaces(&m); if (!vec3(&m, &nmx->m.emit)) break; skip_spaces(&m);
            nm++;
        }
        free(mt);
        s->n_mats = (int)nm; s->mats = (material *)malloc((nm ? nm : 1) * sizeof(material));
        for (size_t i = 0; i < nm; i++) s->mats[i] = mats[i].m;
        /* Obj.hs:73-86 : every (object, material) pair whose names match, object-major */
        size_t cap_t = nf ? nf : 1, nt = 0; s->tris = (triangle *)malloc(cap_t * sizeof(triangle));
        for (size_t oi = 0; oi < no; oi++)
            for (size_t mi = 0; mi < nm; mi++) {
                if (strcmp(objs[oi].mtl, mats[mi].name) != 0) continue;
                for (int fi = 0; fi < objs[oi].nf; fi++) {
                    int *f = &faces[3 * (objs[oi].f0 + fi)];
                    if (f[0] < 1 || f[1] < 1 || f[2] < 1 || (size_t)f[0] > nv || (size_t)f[1] > nv || (size_t)f[2] > nv) {
                        snprintf(s->err, sizeof s->err, "face index out of range"); free(mats); goto done;
                    }
                    if (nt == cap_t) { cap_t *= 2; s->tris = (triangle *)realloc(s->tris, cap_t * sizeof(triangle)); }
                    triangle t = { verts[f[0] - 1], verts[f[1] - 1], verts[f[2] - 1], (int)mi };   /* vs !! (a-1) */
                    s->tris[nt++] = t;
                }
            }
        s->n_tris = (int)nt;
        free(mats);
    }
done:
    free(txt); free(verts); free(faces); free(objs);
    return s;
}

/* Synthetic scenes: triangles as 9 floats each (already in scene space), material index each. */
orc_scene *orc_scene_from_arrays(const float *v9, const int *mat_idx, int n_tris, const float *mats8, int n_mats)
{
    orc_scene *s = (orc_scene *)calloc(1, sizeof *s);
    s->n_tris = n_tris; s->tris = (triangle *)malloc((n_tris ? n_tris : 1) * sizeof(triangle));
    for (int i = 0; i < n_tris; i++) {
        const float *p = v9 + 9 * (size_t)i;
        triangle t = { V(p[0], p[1], p[2]), V(p[3], p[4], p[5]), V(p[6], p[7], p[8]), mat_idx[i] };
        s->tris[i] = t;
    }
    s->n_mats = n_mats; s->mats = (material *)malloc((n_mats ? n_mats : 1) * sizeof(material));
    for (int i = 0; i < n_mats; i++) {
        const float *m = mats8 + 8 * i;
        material mm = { m[0], V(m[1], m[2], m[3]), m[4], V(m[5], m[6], m[7]) };
        s->mats[i] = mm;
    }
    return s;
}

void orc_free_scene(orc_scene *s)
{
    if (!s) return;
    free(s->tris); free(s->mats); free(s->nodes); free(s->leaf_tris); free(s->spheres); free(s);
}
const char *orc_error(orc_scene *s) { return s->err; }
int orc_n_tris(orc_scene *s) { return s->n_tris; }
int orc_n_mats(orc_scene *s) { return s->n_mats; }
int orc_n_nodes(orc_scene *s) { return s->n_nodes; }
int orc_height(orc_scene *s) { return s->height; }
int orc_longest_leaf(orc_scene *s) { return s->longest_leaf; }
int orc_n_leaves(orc_scene *s) { return s->n_leaves; }
void orc_get_tris(orc_scene *s, float *v9, int *mat_idx)
{
    for (int i = 0; i < s->n_tris; i++) {
        triangle *t = &s->tris[i]; float *p = v9 + 9 * (size_t)i;
        p[0] = t->a.x; p[1] = t->a.y; p[2] = t->a.z; p[3] = t->b.x; p[4] = t->b.y; p[5] = t->b.z;
        p[6] = t->c.x; p[7] = t->c.y; p[8] = t->c.z; mat_idx[i] = t->mat;
    }
}
void orc_get_mats(orc_scene *s, float *m8)
{
    for (int i = 0; i < s->n_mats; i++) {
        material *m = &s->mats[i]; float *p = m8 + 8 * i;
        p[0] = m->reflective; p[1] = m->surf.x; p[2] = m->surf.y; p[3] = m->surf.z;
        p[4] = m->emissive; p[5] = m->emit.x; p[6] = m->emit.y; p[7] = m->emit.z;
    }
}

/* ============================================================ Geometry.hs */

/* Geometry.hs:79-80 */
static inline v3 tri_normal(const triangle *t) { return vcross(vsub(t->b, t->a), vsub(t->c, t->a)); }

/* Geometry.hs:90-102 with Data.Matrix multStd: c_ij = sum_k a_ik*b_kj, `sum` = foldl (+) 0.
 * foldr1 (*) [Rz, Ry, Rx] = Rz * (Ry * Rx).  (matrix pkg un-vendored: summation order unpinned.) */
static void mat3mul(const float *a, const float *b, float *c)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float acc = 0.0f;
            for (int k = 0; k < 3; k++) acc = acc + a[3 * i + k] * b[3 * k + j];
            c[3 * i + j] = acc;
        }
}
void orc_rot_matrix_rads(float alp, float bet, float gam, float *R)
{
    float rz[9] = { cosf(alp), -sinf(alp), 0, sinf(alp), cosf(alp), 0, 0, 0, 1 };
    float ry[9] = { cosf(bet), 0, sinf(bet), 0, 1, 0, -sinf(bet), 0, cosf(bet) };
    float rx[9] = { 1, 0, 0, 0, cosf(gam), -sinf(gam), 0, sinf(gam), cosf(gam) };
    float t[9];
    mat3mul(ry, rx, t);
    mat3mul(rz, t, R);
}
/* Geometry.hs:104-107 : row vector times matrix */
static inline v3 rot_vert(v3 v, const float *R)
{
    float in[3] = { v.x, v.y, v.z }, out[3];
    for (int j = 0; j < 3; j++) {
        float acc = 0.0f;
        for (int k = 0; k < 3; k++) acc = acc + in[k] * R[3 * k + j];
        out[j] = acc;
    }
    return V(out[0], out[1], out[2]);
}
/* Obj.hs:60-70 : "px py pz\n a b g" ; cam[0..2] = position, cam[3..11] = rotation row-major */
int orc_load_camera(const char *path, float *cam12)
{
    char *txt = slurp(path); if (!txt) return 1;
    cursor c = { txt }; v3 pos, ang;
    int ok = vec3(&c, &pos) && vec3(&c, &ang);
    free(txt);
    if (!ok) return 2;
    cam12[0] = pos.x; cam12[1] = pos.y; cam12[2] = pos.z;
    orc_rot_matrix_rads(ang.x, ang.y, ang.z, cam12 + 3);
    return 0;
}

/* Geometry.hs:117-142 */
static inline isect moller_trumbore(ray r, const triangle *tri, int tri_index)
{
    isect none = { 0, { 0, 0, 0 }, 0, -1 };
    const float eps = 0.0001f;
    v3 edge1 = vsub(tri->b, tri->a);
    v3 edge2 = vsub(tri->c, tri->a);
    v3 h = vcross(r.d, edge2);
    float a = vdot(edge1, h);
    if (a > -eps && a < eps) return none;
    float f = 1.0f / a;
    v3 s = vsub(r.o, tri->a);
    float u = f * vdot(s, h);
    if (u < 0 || u > 1) return none;
    v3 q = vcross(s, edge1);
    float v = f * vdot(r.d, q);
    if (v < 0 || u + v > 1) return none;
    float t = f * vdot(edge2, q);
    if (t > eps) {
        isect out;
        out.hit = 1;
        out.point = vadd(r.o, vscale(t, r.d));
        out.dist = vnorm(vsub(out.point, r.o));
        out.tri = tri_index;
        return out;
    }
    return none;
}

/* `comparing dist` : LT if a<b, EQ if a==b, else GT (so anything vs NaN is GT) */
static inline int cmp_gt(float a, float b) { return !(a < b) && !(a == b); }

/* Data.Foldable.minimumBy in base-4.9 (GHC 8.0.2) = foldr1 min' where
 * min' x y = case cmp x y of GT -> y ; _ -> x .  Right fold: start from the last element. */
static inline isect min_prime(isect x, isect y) { return cmp_gt(x.dist, y.dist) ? y : x; }

/* Geometry.hs:110-115 */
static isect naive_intersect(const orc_scene *s, ray r, orc_counters *cn)
{
    isect best = { 0, { 0, 0, 0 }, 0, -1 };
    for (int i = s->n_tris - 1; i >= 0; i--) {
        isect h = moller_trumbore(r, &s->tris[i], i);
        if (cn) cn->tri_tests++;
        if (!h.hit) continue;
        best = best.hit ? min_prime(h, best) : h;
    }
    return best;
}

/* Geometry.hs:166-177 */
static inline int intersects_bb(bounds b, ray r)
{
    float dfx = 1.0f / r.d.x, dfy = 1.0f / r.d.y, dfz = 1.0f / r.d.z;
    float t1 = (b.lo.x - r.o.x) * dfx;
    float t2 = (b.hi.x - r.o.x) * dfx;
    float t3 = (b.lo.y - r.o.y) * dfy;
    float t4 = (b.hi.y - r.o.y) * dfy;
    float t5 = (b.lo.z - r.o.z) * dfz;
    float t6 = (b.hi.z - r.o.z) * dfz;
    float tmin = hs_max(hs_max(hs_min(t1, t2), hs_min(t3, t4)), hs_min(t5, t6));
    float tmax = hs_min(hs_min(hs_max(t1, t2), hs_max(t3, t4)), hs_max(t5, t6));
    return tmax > 0 && tmin < tmax;
}

/* Geometry.hs:155-163 over `concatMap vertices` (Geometry.hs:196-197).
 * minimum/maximum on [Float] = strict left fold with the class-default min/max. */
static bounds bounding_box(const orc_scene *s, const int *idx, int n)
{
    bounds b;
    int first = 1;
    for (int i = 0; i < n; i++) {
        const triangle *t = &s->tris[idx[i]];
        v3 vs[3] = { t->a, t->b, t->c };
        for (int k = 0; k < 3; k++) {
            if (first) { b.lo = vs[k]; b.hi = vs[k]; first = 0; continue; }
            b.lo.x = hs_min(b.lo.x, vs[k].x); b.hi.x = hs_max(b.hi.x, vs[k].x);
            b.lo.y = hs_min(b.lo.y, vs[k].y); b.hi.y = hs_max(b.hi.y, vs[k].y);
            b.lo.z = hs_min(b.lo.z, vs[k].z); b.hi.z = hs_max(b.hi.z, vs[k].z);
        }
    }
    if (first) { b.lo = V(0, 0, 0); b.hi = V(0, 0, 0); }   /* reference would throw on an empty list */
    return b;
}

/* Geometry.hs:191-193 : maximumBy (comparing snd) = foldr1 max', max' x y = GT -> x ; _ -> y
 * => ties prefer the later axis (Z over Y over X). */
static int longest_axis(bounds b)
{
    float dx = b.hi.x - b.lo.x, dy = b.hi.y - b.lo.y, dz = b.hi.z - b.lo.z;
    int inner = cmp_gt(dy, dz) ? 1 : 2;
    float iv = inner == 1 ? dy : dz;
    return cmp_gt(dx, iv) ? 0 : inner;
}

/* Geometry.hs:181-182 on `vertices tri`: sum = foldl (+) 0, then each component / 3 */
static inline float centroid_axis(const triangle *t, int ax)
{
    float s = 0.0f;
    s = s + axis_of(t->a, ax);
    s = s + axis_of(t->b, ax);
    s = s + axis_of(t->c, ax);
    return s / 3.0f;
}

/* ================================================================= BIH.hs */
static int new_node(orc_scene *s)
{
    if (s->n_nodes == s->cap_nodes) {
        s->cap_nodes = s->cap_nodes ? s->cap_nodes * 2 : 1024;
        s->nodes = (bnode *)realloc(s->nodes, (size_t)s->cap_nodes * sizeof(bnode));
    }
    memset(&s->nodes[s->n_nodes], 0, sizeof(bnode));
    return s->n_nodes++;
}
static int make_leaf(orc_scene *s, const int *idx, int n)
{
    int id = new_node(s);
    s->nodes[id].leaf = 1; s->nodes[id].first = s->n_leaf_tris; s->nodes[id].count = n;
    memcpy(s->leaf_tris + s->n_leaf_tris, idx, (size_t)n * sizeof(int));
    s->n_leaf_tris += n;
    s->n_leaves++;
    if (n > s->longest_leaf) s->longest_leaf = n;
    return id;
}

/* BIH.hs:67-96.  idx[0..n) is the current triangle list in order; tmp has room for n ints.
 * Nodes are numbered in pre-order; leaves append to leaf_tris left to right, which is
 * exactly `flatten` (BIH.hs:50-52). */
static int bih_build(orc_scene *s, bounds bbox, int *idx, int n, int *tmp, int depth)
{
    if (depth > s->height) s->height = depth;
    if (n < 15) return make_leaf(s, idx, n);                       /* BIH.hs:69,80 */
    /* split, BIH.hs:82-96 */
    int ax = longest_axis(bbox);
    float acc = 0.0f;                                             /* sum = foldl (+) 0 */
    for (int i = 0; i < n; i++) acc = acc + centroid_axis(&s->tris[idx[i]], ax);
    /* genericLength at Float is 1+(1+(...)) in Float arithmetic: saturates at 2^24 */
    float nf = n <= 16777216 ? (float)n : 16777216.0f;
    float split_plane = acc / nf;
    int nl = 0, nr = 0;
    for (int i = 0; i < n; i++) {
        if (centroid_axis(&s->tris[idx[i]], ax) < split_plane) idx[nl++] = idx[i];   /* stable, nl <= i */
        else tmp[nr++] = idx[i];
    }
    memcpy(idx + nl, tmp, (size_t)nr * sizeof(int));
    int *left = idx, *right = idx + nl;
    float lbest = axis_of(bbox.lo, ax), rbest = axis_of(bbox.hi, ax);    /* maximumDef leftSide / minimumDef rightSide */
    for (int i = 0; i < nl; i++) {
        const triangle *t = &s->tris[left[i]];
        float c0 = axis_of(t->a, ax), c1 = axis_of(t->b, ax), c2 = axis_of(t->c, ax);
        if (i == 0) lbest = c0; else lbest = hs_max(lbest, c0);
        lbest = hs_max(lbest, c1); lbest = hs_max(lbest, c2);
    }
    for (int i = 0; i < nr; i++) {
        const triangle *t = &s->tris[right[i]];
        float c0 = axis_of(t->a, ax), c1 = axis_of(t->b, ax), c2 = axis_of(t->c, ax);
        if (i == 0) rbest = c0; else rbest = hs_min(rbest, c0);
        rbest = hs_min(rbest, c1); rbest = hs_min(rbest, c2);
    }
    float lmax = 0.001f + lbest;
    float rmin = (-0.001f) + rbest;
    int id = new_node(s);
    s->nodes[id].leaf = 0; s->nodes[id].axis = ax; s->nodes[id].lmax = lmax; s->nodes[id].rmin = rmin;
    int l, r;
    if (nl == 0) {                                                /* BIH.hs:70-72 */
        if (depth + 1 > s->height) s->height = depth + 1;
        l = make_leaf(s, left, 0); r = make_leaf(s, right, nr);
    } else if (nr == 0) {                                         /* BIH.hs:73-75 */
        if (depth + 1 > s->height) s->height = depth + 1;
        l = make_leaf(s, left, nl); r = make_leaf(s, right, 0);
    } else {                                                      /* BIH.hs:76-78 */
        l = bih_build(s, bounding_box(s, left, nl), left, nl, tmp, depth + 1);
        r = bih_build(s, bounding_box(s, right, nr), right, nr, tmp, depth + 1);
    }
    s->nodes[id].left = l; s->nodes[id].right = r;
    return id;
}

/* BIH.hs:62-65 */
int orc_make_bih(orc_scene *s)
{
    free(s->nodes); free(s->leaf_tris);
    s->nodes = NULL; s->n_nodes = s->cap_nodes = 0; s->n_leaf_tris = 0;
    s->height = 0; s->longest_leaf = 0; s->n_leaves = 0;
    int n = s->n_tris;
    s->leaf_tris = (int *)malloc((size_t)(n ? n : 1) * sizeof(int));
    int *idx = (int *)malloc((size_t)(n ? n : 1) * sizeof(int));
    int *tmp = (int *)malloc((size_t)(n ? n : 1) * sizeof(int));
    for (int i = 0; i < n; i++) idx[i] = i;
    s->root = bounding_box(s, idx, n);
    bih_build(s, s->root, idx, n, tmp, 1);
    free(idx); free(tmp);
    return 0;
}

/* Flattened export for cross-checking the product host's flatten: per node 4 x u32
 * {lmax bits, rmin bits, a, b}; branch: a = left | axis<<30, b = right ; leaf: a = first, b = count | 0x80000000 */
void orc_export_bih(orc_scene *s, float *root6, uint32_t *nodes4, int *leaf_tris)
{
    root6[0] = s->root.lo.x; root6[1] = s->root.lo.y; root6[2] = s->root.lo.z;
    root6[3] = s->root.hi.x; root6[4] = s->root.hi.y; root6[5] = s->root.hi.z;
    for (int i = 0; i < s->n_nodes; i++) {
        bnode *n = &s->nodes[i]; uint32_t *o = nodes4 + 4 * (size_t)i;
        if (n->leaf) { o[0] = 0; o[1] = 0; o[2] = (uint32_t)n->first; o[3] = (uint32_t)n->count | 0x80000000u; }
        else {
            memcpy(&o[0], &n->lmax, 4); memcpy(&o[1], &n->rmin, 4);
            o[2] = (uint32_t)n->left | ((uint32_t)n->axis << 30); o[3] = (uint32_t)n->right;
        }
    }
    memcpy(leaf_tris, s->leaf_tris, (size_t)s->n_leaf_tris * sizeof(int));
}

/* BIH.hs:104-141, literal (including the redundant own-box test at :112) */
static isect intersect_bih_rec(const orc_scene *s, bounds bbox, int node, ray r, orc_counters *cn)
{
    isect none = { 0, { 0, 0, 0 }, 0, -1 };
    const bnode *n = &s->nodes[node];
    if (n->leaf) {                                                /* BIH.hs:105-109 */
        isect best = none;
        for (int i = n->count - 1; i >= 0; i--) {                 /* foldr1 min' */
            int ti = s->leaf_tris[n->first + i];
            isect h = moller_trumbore(r, &s->tris[ti], ti);
            if (cn) cn->tri_tests++;
            if (!h.hit) continue;
            best = best.hit ? min_prime(h, best) : h;
        }
        return best;
    }
    if (cn) cn->own_box_tests++;
    if (!intersects_bb(bbox, r)) return none;                     /* BIH.hs:112 */
    int ax = n->axis;
    bounds left = bbox, right = bbox;                             /* BIH.hs:130-141 */
    if (ax == 0) { left.hi.x = n->lmax; right.lo.x = n->rmin; }
    else if (ax == 1) { left.hi.y = n->lmax; right.lo.y = n->rmin; }
    else { left.hi.z = n->lmax; right.lo.z = n->rmin; }
    int il = intersects_bb(left, r), ir = intersects_bb(right, r);
    if (cn) { cn->branch_visits++; cn->child_box_tests += 2; }
    int left_to_right = axis_of(r.d, ax) > 0;                     /* BIH.hs:127 */
    if (il && ir) {                                               /* BIH.hs:113-116 */
        isect near = left_to_right ? intersect_bih_rec(s, left, n->left, r, cn)
                                   : intersect_bih_rec(s, right, n->right, r, cn);
        if (near.hit) {
            float p = axis_of(near.point, ax);
            int close = left_to_right ? (p < n->rmin) : (p > n->lmax);     /* BIH.hs:121-123 */
            if (close) return near;
        }
        isect far = left_to_right ? intersect_bih_rec(s, right, n->right, r, cn)
                                  : intersect_bih_rec(s, left, n->left, r, cn);
        if (near.hit) return far.hit ? min_prime(near, far) : near;        /* minimumByMay over catMaybes [near, far] */
        return far;
    }
    if (il) return intersect_bih_rec(s, left, n->left, r, cn);    /* BIH.hs:117 */
    if (ir) return intersect_bih_rec(s, right, n->right, r, cn);  /* BIH.hs:118 */
    return none;
}
/* EXTENSION (the reference has no sphere): double-sided analytic sphere with the conventions of mollerTrumbore --
 * eps = 1e-4, point = o + t *^ d, dist = norm (point - o).  Restates the semantics fixed in include/sqt.h. */
static inline isect ray_sphere(ray r, const sphere *sp, int surface_index)
{
    isect none = { 0, { 0, 0, 0 }, 0, -1 };
    const float eps = 0.0001f;
    v3 oc = vsub(r.o, sp->c);
    float A = vdot(r.d, r.d);
    float B = vdot(oc, r.d);
    float C = vdot(oc, oc) - sp->r * sp->r;
    float disc = B * B - A * C;
    if (!(disc >= 0)) return none;
    float sq = sqrtf(disc);
    float t0 = (-B - sq) / A, t1 = (-B + sq) / A;
    float t = t0 > eps ? t0 : t1;
    if (!(t > eps)) return none;
    isect out;
    out.hit = 1;
    out.point = vadd(r.o, vscale(t, r.d));
    out.dist = vnorm(vsub(out.point, r.o));
    out.tri = surface_index;
    return out;
}
/* candidates in order [accelerated hit, sphere 0, sphere 1, ...], minimumBy (comparing dist): the first minimal wins */
static inline isect with_spheres(const orc_scene *s, ray r, isect best)
{
    for (int k = 0; k < s->n_spheres; k++) {
        isect h = ray_sphere(r, &s->spheres[k], s->n_tris + k);
        if (!h.hit) continue;
        if (!best.hit || cmp_gt(best.dist, h.dist)) best = h;
    }
    return best;
}
/* BIH.hs:101-102 (+ the sphere extension) */
static inline isect intersect_bih(const orc_scene *s, ray r, orc_counters *cn)
{
    if (cn) cn->rays++;
    return with_spheres(s, r, intersect_bih_rec(s, s->root, 0, r, cn));
}
/* material and un-normalised geometric normal of the surface an intersection lies on */
static inline const material *surface_material(const orc_scene *s, const isect *in)
{
    return in->tri >= s->n_tris ? &s->mats[s->spheres[in->tri - s->n_tris].mat] : &s->mats[s->tris[in->tri].mat];
}
static inline v3 surface_normal(const orc_scene *s, const isect *in)
{
    if (in->tri >= s->n_tris) return vsub(in->point, s->spheres[in->tri - s->n_tris].c);
    return tri_normal(&s->tris[in->tri]);
}
void orc_set_spheres(orc_scene *s, const float *s5, int n)      /* rows (cx, cy, cz, r, material index) */
{
    free(s->spheres);
    s->spheres = (sphere *)malloc((size_t)(n ? n : 1) * sizeof(sphere));
    s->n_spheres = n;
    for (int k = 0; k < n; k++) {
        sphere sp = { V(s5[5 * k], s5[5 * k + 1], s5[5 * k + 2]), s5[5 * k + 3], (int)s5[5 * k + 4] };
        s->spheres[k] = sp;
    }
}

/* ============================================================ thread pool */
typedef void (*chunk_fn)(void *ctx, int64_t lo, int64_t hi, int tid);
typedef struct { chunk_fn fn; void *ctx; int64_t n, chunk; int64_t *next; int tid; } worker_arg;
static void *worker(void *p)
{
    worker_arg *w = (worker_arg *)p;
    for (;;) {
        int64_t lo = __atomic_fetch_add(w->next, w->chunk, __ATOMIC_RELAXED);
        if (lo >= w->n) break;
        int64_t hi = lo + w->chunk < w->n ? lo + w->chunk : w->n;
        w->fn(w->ctx, lo, hi, w->tid);
    }
    return NULL;
}
static void parallel_for(int64_t n, int64_t chunk, int nthreads, chunk_fn fn, void *ctx)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 512) nthreads = 512;
    int64_t next = 0;
    pthread_t th[512]; worker_arg wa[512];
    for (int t = 0; t < nthreads; t++) {
        worker_arg a = { fn, ctx, n, chunk, &next, t }; wa[t] = a;
        if (t > 0) pthread_create(&th[t], NULL, worker, &wa[t]);
    }
    worker(&wa[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* ======================================================= batched intersect */
typedef struct {
    const orc_scene *s; const float *org, *dir; int naive;
    int32_t *tri_out; float *dist_out, *point_out; orc_counters *cn; /* per thread */
} batch_ctx;
static void batch_chunk(void *p, int64_t lo, int64_t hi, int tid)
{
    batch_ctx *b = (batch_ctx *)p;
    orc_counters *cn = b->cn ? &b->cn[tid] : NULL;
    for (int64_t i = lo; i < hi; i++) {
        ray r = { V(b->org[3 * i], b->org[3 * i + 1], b->org[3 * i + 2]), V(b->dir[3 * i], b->dir[3 * i + 1], b->dir[3 * i + 2]) };
        isect h;
        if (b->naive) { if (cn) cn->rays++; h = with_spheres(b->s, r, naive_intersect(b->s, r, cn)); }
        else h = intersect_bih(b->s, r, cn);
        b->tri_out[i] = h.hit ? h.tri : -1;
        if (b->dist_out) b->dist_out[i] = h.hit ? h.dist : 0.0f;
        if (b->point_out) {
            b->point_out[3 * i] = h.hit ? h.point.x : 0.0f;
            b->point_out[3 * i + 1] = h.hit ? h.point.y : 0.0f;
            b->point_out[3 * i + 2] = h.hit ? h.point.z : 0.0f;
        }
    }
}
/* Scene.intersect (Geometry.hs:62-65) over a batch; naive=1 selects naiveIntersect (Main.hs:52-53).
 * org/dir xyz-interleaved.  counters5 (optional) = {branch_visits, child_box_tests, own_box_tests, tri_tests, rays}. */
int orc_intersect_batch(orc_scene *s, const float *org, const float *dir, int64_t n, int naive,
                        int32_t *tri_out, float *dist_out, float *point_out, uint64_t *counters5, int nthreads)
{
    if (!naive && s->n_nodes == 0) return 1;
    if (nthreads < 1) nthreads = 1;
    orc_counters *cn = counters5 ? (orc_counters *)calloc((size_t)nthreads, sizeof *cn) : NULL;
    batch_ctx b = { s, org, dir, naive, tri_out, dist_out, point_out, cn };
    parallel_for(n, 1024, nthreads, batch_chunk, &b);
    if (cn) {
        memset(counters5, 0, 5 * sizeof(uint64_t));
        for (int t = 0; t < nthreads; t++) {
            counters5[0] += cn[t].branch_visits; counters5[1] += cn[t].child_box_tests;
            counters5[2] += cn[t].own_box_tests; counters5[3] += cn[t].tri_tests; counters5[4] += cn[t].rays;
        }
        free(cn);
    }
    return 0;
}

/* ==================================================================== RNG */
/* Philox4x32-10 (Salmon et al., SC'11).  Replaces tf-random's TFGen: stream = the integer the
 * reference passes to mkTFGen (Lib.hs:85-86), draw = index of the Word32 in that generator's
 * output sequence. */
static inline void philox_round(uint32_t *c, uint32_t k0, uint32_t k1)
{
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
void orc_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out)
{
    uint32_t c[4] = { ctr[0], ctr[1], ctr[2], ctr[3] }, k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; i++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof c);
}
/* draw `j` of stream `stream` under `seed` : counter = (stream lo, stream hi, j/4, "SQTR"), key = seed */
static inline uint32_t draw_word(uint64_t seed, uint64_t stream, uint32_t j)
{
    uint32_t ctr[4] = { (uint32_t)stream, (uint32_t)(stream >> 32), j >> 2, 0x52545153u };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) }, out[4];
    orc_philox4x32_10(ctr, key, out);
    return out[j & 3];
}
/* Lib.hs:183-188 with (lo,hi) = (0,1): p = fromIntegral n / fromIntegral (maxBound::Word32),
 * both converted to Float (denominator rounds to 2^32); result 0 + (1-0)*p.  Inclusive of 1.0. */
static inline float random_r01(uint32_t n)
{
    float p = (float)n / 4294967296.0f;
    float r = 1.0f - 0.0f;
    return 0.0f + r * p;
}

/* =================================================================== trig */
/* "sqt trig" (DESIGN.md): plain binary32, round-to-nearest, no fusion, exactly this order. */
static void sqt_sincos(float x, float *s_out, float *c_out)
{
    /* x in [0, 2pi]; k = nearest multiple of pi/2 */
    int k = (int)(x * 0.63661975f + 0.5f);
    float kf = (float)k;
    float r = ((x - kf * 1.5703125f) - kf * 4.837512969970703125e-4f) - kf * 7.54978995489188216e-8f;
    float z = r * r;
    float s = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float c = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z;
    c = c - 0.5f * z;
    c = c + 1.0f;
    switch (k & 3) {
    case 0: *s_out = s; *c_out = c; break;
    case 1: *s_out = c; *c_out = -s; break;
    case 2: *s_out = -s; *c_out = -c; break;
    default: *s_out = -c; *c_out = s; break;
    }
}
static float sqt_asin_core(float x, float z)
{
    return ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z
            + 1.6666752422e-1f) * z * x + x;
}
static float sqt_acos(float x)
{
    if (x < -0.5f) {
        float z = 0.5f * (1.0f + x); float y = sqrtf(z);
        return 3.14159265358979323846f - 2.0f * sqt_asin_core(y, z);
    }
    if (x > 0.5f) {
        float z = 0.5f * (1.0f - x); float y = sqrtf(z);
        return 2.0f * sqt_asin_core(y, z);
    }
    return 1.57079632679489661923f - sqt_asin_core(x, x * x);
}
static float sqt_atan(float x)      /* x >= 0 */
{
    float y;
    if (x > 2.414213562373095f) { y = 1.57079632679489661923f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = 0.78539816339744830962f; x = (x - 1.0f) / (x + 1.0f); }
    else y = 0.0f;
    float z = x * x;
    y = y + ((((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * x + x);
    return y;
}

/* ================================================================= Lib.hs */
typedef struct {
    int32_t rows, cols;        /* array extent: massiv Ix2 = (row :. col) */
    int32_t xdiv, ydiv;        /* makeRay divisors: reference-literal = (w, h) */
    int32_t seed_stride;       /* rix = spp * (x + y*seed_stride), reference-literal = w */
    int32_t spp, max_depth;    /* reference: max_depth = 3 (Lib.hs:129) */
    int32_t mode;              /* 0 = raytrace, 1 = raycast (--cast) */
    int32_t trig;              /* 0 = libm (as GHC), 1 = sqt trig */
    int32_t rank, world;       /* pixel-group partition (group = 32 consecutive pixels, round-robin); world<=1: all */
    int32_t split_samples;     /* 1: partition by sample range instead of pixel groups */
    uint64_t seed;
} orc_params;

typedef struct {
    const orc_scene *s; const float *cam; orc_params p;
    float *accum; uint8_t *rgb8; orc_counters *cn; uint64_t *samples;
} render_ctx;

/* Lib.hs:107-114 (pixel row index y, column index x) */
static inline ray make_ray(const orc_params *p, int y, int x, const float *cam)
{
    float ww = (float)p->xdiv, hh = (float)p->ydiv;
    float xoffs = ((float)x - (ww / 2)) / ww;
    float yoffs = ((hh / 2) - (float)y) / hh;
    ray r = { V(cam[0], cam[1], cam[2]), rot_vert(V(1, xoffs, yoffs), cam + 3) };
    return r;
}
void orc_make_rays(const orc_params *p, const float *cam, float *org, float *dir)
{
    for (int y = 0; y < p->rows; y++)
        for (int x = 0; x < p->cols; x++) {
            ray r = make_ray(p, y, x, cam); size_t i = (size_t)y * p->cols + x;
            org[3 * i] = r.o.x; org[3 * i + 1] = r.o.y; org[3 * i + 2] = r.o.z;
            dir[3 * i] = r.d.x; dir[3 * i + 1] = r.d.y; dir[3 * i + 2] = r.d.z;
        }
}

/* Lib.hs:192-198 ; u is draw j, v is draw j+1 of the sample's stream */
static v3 random_vector(float u, float v, int trig)
{
    float th = 2 * 3.14159265358979323846f * u;      /* 2 * pi * u, left-assoc: (2*pi) folded in Float */
    float sth, cth, sph, cph;
    if (trig == 0) {
        float ph = acosf(2 * v - 1);
        cth = cosf(th); sth = sinf(th); sph = sinf(ph); cph = cosf(ph);
    } else {
        float ph = sqt_acos(2 * v - 1);
        sqt_sincos(th, &sth, &cth); sqt_sincos(ph, &sph, &cph);
    }
    return V(cth * sph, sth * sph, cph);
}

/* Lib.hs:127-137 with bounceRay/scatterRay/reflectRay (Lib.hs:155-181) inlined in order */
static v3 raytrace(const render_ctx *rc, uint64_t stream, ray r, int bounces, orc_counters *cn)
{
    const v3 black = V(0, 0, 0);
    if (bounces > rc->p.max_depth - 1) return black;               /* reference: bounces > 2 */
    isect in = intersect_bih(rc->s, r, cn);
    if (!in.hit) return black;
    const material *m = surface_material(rc->s, &in);
    v3 next = black;
    if (bounces + 1 <= rc->p.max_depth - 1) {                      /* lazy: newRay only forced if traced */
        float x = random_r01(draw_word(rc->p.seed, stream, (uint32_t)bounces));
        ray nr;
        if (m->reflective < x) {                                   /* Lib.hs:157 scatterRay, Lib.hs:166-172 */
            float v = random_r01(draw_word(rc->p.seed, stream, (uint32_t)bounces + 1));
            v3 nd = random_vector(x, v, rc->p.trig);
            v3 n = surface_normal(rc->s, &in);
            float old = hs_signum(vdot(r.d, n));
            float nw = hs_signum(vdot(nd, n));
            nr.o = in.point;
            nr.d = (old == nw) ? vneg(nd) : nd;
        } else {                                                   /* Lib.hs:176-181 */
            v3 dn = vnormalize(surface_normal(rc->s, &in));
            v3 di = r.d;
            nr.o = in.point;
            nr.d = vsub(di, vscale(2 * vdot(dn, di), dn));
        }
        next = raytrace(rc, stream, nr, bounces + 1, cn);          /* newGen = snd (next gen): draw index + 1 */
    }
    v3 next_bounce = vmul(m->surf, next);
    v3 emit = vscale(m->emissive, m->emit);
    return vadd(next_bounce, emit);
}

/* Lib.hs:141-151 */
static v3 raycast(const render_ctx *rc, ray r, orc_counters *cn)
{
    const v3 black = V(0, 0, 0);
    isect in = intersect_bih(rc->s, r, cn);
    if (!in.hit) return black;
    const material *m = surface_material(rc->s, &in);
    v3 light = V(0, 3, -1);
    ray shadow = { in.point, vsub(light, in.point) };              /* a `to` b = Ray a (b - a) */
    float dl = vnorm(vsub(in.point, light));
    isect sh = intersect_bih(rc->s, shadow, cn);
    if (sh.hit && !(sh.dist > dl)) return black;                   /* guard $ maybe True (\pos -> dist pos > dl) */
    return vscale(2 / dl, m->surf);
}

/* Lib.hs:93-104.  floor :: Float -> Word8 goes through Integer and wraps mod 256; NaN -> 0. */
static inline uint8_t to_w8(float s255)
{
    if (s255 != s255) return 0;
    float f = floorf(s255);
    /* |f| >= 2^32 (or inf): the Integer is a multiple of 256 -> 0 */
    int64_t w = fabsf(f) < 9.0e18f ? (int64_t)f : 0;
    uint32_t b = (uint32_t)((uint64_t)w & 0xff);
    return (uint8_t)(b < 255 ? b : 255);
}
static void tone_map(v3 c, int trig, uint8_t *out)
{
    float maxc = hs_max(hs_max(c.x, c.y), c.z);
    float minc = hs_min(hs_min(c.x, c.y), c.z);
    float lightness = 0.5f * (maxc + minc);
    float at = trig == 0 ? atanf(lightness) : sqt_atan(lightness);
    float intensity = at / (3.14159265358979323846f / 2);
    v3 s1 = vscale(intensity / maxc, c);
    out[0] = to_w8(s1.x * 255); out[1] = to_w8(s1.y * 255); out[2] = to_w8(s1.z * 255);
}
void orc_tone_map(const float *accum, int64_t n_pixels, int spp, int trig, uint8_t *rgb8)
{
    float inv = 1 / (float)spp;
    for (int64_t i = 0; i < n_pixels; i++)
        tone_map(vscale(inv, V(accum[3 * i], accum[3 * i + 1], accum[3 * i + 2])), trig, rgb8 + 3 * i);
}

static inline int owns_pixel(const orc_params *p, int64_t pix)
{
    if (p->world <= 1 || p->split_samples) return 1;
    return (int)((pix >> 5) % p->world) == p->rank;
}

/* Lib.hs:79-89 */
static void render_chunk(void *ctx, int64_t lo, int64_t hi, int tid)
{
    render_ctx *rc = (render_ctx *)ctx;
    const orc_params *p = &rc->p;
    orc_counters *cn = &rc->cn[tid];
    int k0 = 0, k1 = p->spp;
    if (p->world > 1 && p->split_samples) {
        k0 = (int)((int64_t)p->spp * p->rank / p->world);
        k1 = (int)((int64_t)p->spp * (p->rank + 1) / p->world);
    }
    for (int64_t pix = lo; pix < hi; pix++) {
        int y = (int)(pix / p->cols), x = (int)(pix % p->cols);
        v3 sum = V(0, 0, 0);                                       /* sum = foldl (+) (fromInteger 0) */
        if (owns_pixel(p, pix)) {
            ray r = make_ray(p, y, x, rc->cam);
            uint64_t rix = (uint64_t)p->spp * ((uint64_t)x + (uint64_t)y * (uint64_t)p->seed_stride);
            for (int k = k0; k < k1; k++) {
                v3 c = p->mode == 1 ? raycast(rc, r, cn) : raytrace(rc, rix + (uint64_t)k, r, 0, cn);
                sum = vadd(sum, c);
                rc->samples[tid]++;
            }
        }
        if (rc->accum) { rc->accum[3 * pix] = sum.x; rc->accum[3 * pix + 1] = sum.y; rc->accum[3 * pix + 2] = sum.z; }
        if (rc->rgb8) tone_map(vscale(1 / (float)p->spp, sum), p->trig, rc->rgb8 + 3 * pix);
    }
}

/* Lib.hs:68-75 minus writeImage.  accum = per-pixel radiance SUM (not yet divided by spp);
 * rgb8 = tone-mapped mean.  stats2 = {rays (isect calls), samples}.  counters5 as above. */
int orc_render(orc_scene *s, const float *cam12, const orc_params *p, float *accum, uint8_t *rgb8,
               uint64_t *stats2, uint64_t *counters5, int nthreads)
{
    if (s->n_nodes == 0) return 1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 512) nthreads = 512;
    render_ctx rc; rc.s = s; rc.cam = cam12; rc.p = *p; rc.accum = accum; rc.rgb8 = rgb8;
    rc.cn = (orc_counters *)calloc((size_t)nthreads, sizeof(orc_counters));
    rc.samples = (uint64_t *)calloc((size_t)nthreads, sizeof(uint64_t));
    parallel_for((int64_t)p->rows * p->cols, 256, nthreads, render_chunk, &rc);
    uint64_t rays = 0, smp = 0; orc_counters tot; memset(&tot, 0, sizeof tot);
    for (int t = 0; t < nthreads; t++) {
        rays += rc.cn[t].rays; smp += rc.samples[t];
        tot.branch_visits += rc.cn[t].branch_visits; tot.child_box_tests += rc.cn[t].child_box_tests;
        tot.own_box_tests += rc.cn[t].own_box_tests; tot.tri_tests += rc.cn[t].tri_tests; tot.rays += rc.cn[t].rays;
    }
    if (stats2) { stats2[0] = rays; stats2[1] = smp; }
    if (counters5) {
        counters5[0] = tot.branch_visits; counters5[1] = tot.child_box_tests; counters5[2] = tot.own_box_tests;
        counters5[3] = tot.tri_tests; counters5[4] = tot.rays;
    }
    free(rc.cn); free(rc.samples);
    return 0;
}

/* Lib.hs:79-89 for the pixels [pix_lo, pix_hi) of a frame only (same seeds and rays as in the full frame), so
 * that tests can check rows of a full-size GPU render.  accum has (pix_hi - pix_lo) * 3 floats. */
typedef struct { render_ctx rc; int64_t off; } window_ctx;
static void window_chunk(void *ctx, int64_t lo, int64_t hi, int tid)
{
    window_ctx *w = (window_ctx *)ctx;
    render_chunk(&w->rc, lo + w->off, hi + w->off, tid);
}
int orc_render_window(orc_scene *s, const float *cam12, const orc_params *p, int64_t pix_lo, int64_t pix_hi, float *accum,
                      int nthreads)
{
    if (s->n_nodes == 0 || pix_lo < 0 || pix_hi > (int64_t)p->rows * p->cols || pix_lo > pix_hi) return 1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 512) nthreads = 512;
    window_ctx w;
    w.rc.s = s; w.rc.cam = cam12; w.rc.p = *p; w.rc.rgb8 = NULL;
    w.rc.accum = accum - 3 * pix_lo;          /* render_chunk indexes by absolute pixel */
    w.rc.cn = (orc_counters *)calloc((size_t)nthreads, sizeof(orc_counters));
    w.rc.samples = (uint64_t *)calloc((size_t)nthreads, sizeof(uint64_t));
    w.off = pix_lo;
    parallel_for(pix_hi - pix_lo, 64, nthreads, window_chunk, &w);
    free(w.rc.cn); free(w.rc.samples);
    return 0;
}

/* helpers exported for unit tests */
float orc_random_r01(uint32_t n) { return random_r01(n); }
uint32_t orc_draw_word(uint64_t seed, uint64_t stream, uint32_t j) { return draw_word(seed, stream, j); }
void orc_sqt_sincos(float x, float *s, float *c) { sqt_sincos(x, s, c); }
float orc_sqt_acos(float x) { return sqt_acos(x); }
float orc_sqt_atan(float x) { return sqt_atan(x); }
int orc_intersects_bb(const float *b6, const float *o3, const float *d3)
{
    bounds b = { V(b6[0], b6[1], b6[2]), V(b6[3], b6[4], b6[5]) };
    ray r = { V(o3[0], o3[1], o3[2]), V(d3[0], d3[1], d3[2]) };
    return intersects_bb(b, r);
}
int orc_moller_trumbore(const float *tri9, const float *o3, const float *d3, float *point3, float *dist)
{
    triangle t = { V(tri9[0], tri9[1], tri9[2]), V(tri9[3], tri9[4], tri9[5]), V(tri9[6], tri9[7], tri9[8]), 0 };
    ray r = { V(o3[0], o3[1], o3[2]), V(d3[0], d3[1], d3[2]) };
    isect h = moller_trumbore(r, &t, 0);
    if (h.hit) { point3[0] = h.point.x; point3[1] = h.point.y; point3[2] = h.point.z; *dist = h.dist; }
    return h.hit;
}
