// oracle/xnart_oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17, strict IEEE binary32 per operator, no FMA contraction) of the
// per-pixel trace loop of eitan3/xna-ray-trace.  It exists to CHECK libxnart.so; nothing in
// the product path may include, link or execute it (only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs do).
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path
//  and its C# cannot be compiled or run in this environment (no mono/dotnet,
// closed-source XNA 4.0), so this restatement is pinned only by hand-computed
// known-answer tests (tests/test_oracle_kat.py) and by the cross-check of its two traversal
// implementations (literal vs pruned) against each other.
//
// Reference lines followed (paths relative to /root/reference):
//   ray generation      RayTraceProject/RayTraceProject/RayTracer.cs:391-429
//   multisampling       RayTraceProject/RayTraceProject/RayTracer.cs:128-311
//   CastRay / shadows   RayTraceProject/RayTraceProject/RayTracer.cs:465-502, 506-737
//   scene octree        RayTraceProject/RayTraceProject/Spatial/OctreeSpatialManager.cs:64-99, 218-248, 312-482
//   mesh octree         RayTracerTypeLibrary/MeshOctree.cs:56-96, 204-236, 259-353
//   mesh prefilter      RayTracerTypeLibrary/Mesh.cs:34-39
//   triangle test       RayTracerTypeLibrary/RayExtensions.cs:42-75
//   texture lookup      RayTracerTypeLibrary/Material.cs:71-232
//   lights              RayTraceProject/RayTraceProject/SpotLight.cs:19-62, DirectionalLight.cs:23-30
//   object transforms   RayTraceProject/RayTraceProject/SceneObject.cs:183-199
// Framework math (closed source) is restated in xna_math.h.
#include "xna_math.h"
#include "../include/xnart.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace xo {

// ------------------------------------------------------------------------------------------
// Scene records
// ------------------------------------------------------------------------------------------
struct Tri {
    V3 v1, v2, v3;
    V2 uv1, uv2, uv3;
    V3 n1, n2, n3;
    V3 sn;
    V4 color;
    int id;
};

struct Material {
    float reflectiveness = 0;
    bool useTexture = false, interpolateNormals = false, transparent = false;
    float refractionIndex = 0;
    int texW = 0, texH = 0;
    std::vector<uint32_t> argb;
};

// One CubeNode (MeshOctree.cs:32-40 / OctreeSpatialManager.cs:37-48), flattened.
struct Node {
    Box b;
    Box cons;          // union of this box and every descendant box (pruned traversal only)
    int firstChild;    // index of child 0 (children are 8 consecutive nodes), -1 for a leaf
    int refStart;      // leaves: offset into refs[], in DFS leaf order
    int refCount;      // number of items in containingObjects
};

struct Tree {
    std::vector<Node> nodes;
    std::vector<int> refs;  // leaf item lists, concatenated in DFS (child 0..7) leaf order
    int depth = 0;
};

struct Mesh {
    std::vector<Tri> tris;
    Material mat;
    Box box;  // Mesh.MeshBoundingBox
    Tree tree;
};

struct Obj {
    std::vector<int> meshes;
    M4 world, inv;
    Box localBox;
    Box worldBox;  // SceneObject.WorldBoundingBox: corners transformed, NOT re-sorted (Q8)
};

struct Light {
    int kind;
    V3 position, direction, color;
    float intensity, decayExponent, spotAngle;
    float angleCosine;  // SpotLight.cs:25
};

struct Counters {
    std::atomic<uint64_t> raysPrimary{0}, raysShadow{0}, raysReflection{0}, raysRefraction{0};
    std::atomic<uint64_t> hits{0}, nBox{0}, nTriRef{0}, nTriCulled{0}, nTriFull{0}, pixels{0};
};

struct LocalCounters {
    uint64_t raysPrimary = 0, raysShadow = 0, raysReflection = 0, raysRefraction = 0;
    uint64_t hits = 0, nBox = 0, nTriRef = 0, nTriCulled = 0, nTriFull = 0, pixels = 0;
};

struct Hit {
    int obj = -1, mesh = -1, tri = -1;
    float u = 0, v = 0, d = 0;
    V3 worldPos{0, 0, 0};
};

enum { MODE_LITERAL = 0, MODE_PRUNED = 1 };

struct Ctx {
    std::vector<std::unique_ptr<Mesh>> meshes;
    std::vector<Obj> objs;
    Tree sceneTree;
    bool built = false;
    std::vector<Light> lights;
    int semantics = XNART_SEMANTICS_XNA40;
    int mode = MODE_LITERAL;
    Counters counters;
    std::string err;
};

static const int kMaxDepth = 64;

// ------------------------------------------------------------------------------------------
// Slab test dispatch (BoundingBox.Intersects(ref Ray, out float?))
// ------------------------------------------------------------------------------------------
static inline bool slab(const Box& b, const Ray& r, float& t, int semantics) {
    return semantics == XNART_SEMANTICS_MONOGAME ? slabMonoGame(b, r, t) : slabXna40(b, r, t);
}

// ------------------------------------------------------------------------------------------
// Octree construction
// ------------------------------------------------------------------------------------------
// MeshOctree.SplitCuboid (MeshOctree.cs:204-236) / OctreeSpatialManager.SplitCuboid (:218-248):
// child index = 4i+2j+k; box = (Min + size*(i,j,k), that + size), size = (Max-Min)/2f.
static void childBoxes(const Box& parent, Box out[8]) {
    V3 size = divs(sub(parent.mx, parent.mn), 2.0f);
    int idx = 0;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) {
                V3 pos = add(parent.mn, V3{size.x * (float)i, size.y * (float)j, size.z * (float)k});
                out[idx++] = Box{pos, add(pos, size)};
            }
}

// BuildTree (MeshOctree.cs:84-96): split while Count > threshold, children visited 0..7.
// `contains(item, box)` is the per-level membership rule.
template <class Contains>
static bool buildRec(Tree& t, int nodeIdx, std::vector<int>& items, int threshold, int depth,
                     const Contains& contains) {
    if ((int)items.size() > threshold) {
        if (depth >= kMaxDepth) return false;
        if (depth + 1 > t.depth) t.depth = depth + 1;
        Box cb[8];
        childBoxes(t.nodes[nodeIdx].b, cb);
        int first = (int)t.nodes.size();
        t.nodes[nodeIdx].firstChild = first;
        std::vector<std::vector<int>> lists(8);
        for (int c = 0; c < 8; c++) {
            Node n;
            n.b = cb[c];
            n.cons = cb[c];
            n.firstChild = -1;
            n.refStart = 0;
            n.refCount = 0;
            t.nodes.push_back(n);
            for (int it : items)
                if (contains(it, cb[c])) lists[c].push_back(it);
        }
        // inner nodes keep containingObjects in the reference but traversal never reads it
        t.nodes[nodeIdx].refCount = (int)items.size();
        std::vector<int>().swap(items);
        for (int c = 0; c < 8; c++)
            if (!buildRec(t, first + c, lists[c], threshold, depth + 1, contains)) return false;
        Box cons = t.nodes[nodeIdx].b;
        for (int c = 0; c < 8; c++) cons = boxMerge(cons, t.nodes[first + c].cons);
        t.nodes[nodeIdx].cons = cons;
    } else {
        Node& n = t.nodes[nodeIdx];
        n.refStart = (int)t.refs.size();
        n.refCount = (int)items.size();
        t.refs.insert(t.refs.end(), items.begin(), items.end());
    }
    return true;
}

// MeshOctree.Build (MeshOctree.cs:56-82): tight root box, then BuildTree.
static bool buildMeshTree(Mesh& m, int threshold) {
    Box box{V3{FLT_MAX, FLT_MAX, FLT_MAX}, V3{-FLT_MAX, -FLT_MAX, -FLT_MAX}};  // float.MinValue == -MaxValue
    for (const Tri& t : m.tris) {
        box.mn = vmin(box.mn, t.v1); box.mn = vmin(box.mn, t.v2); box.mn = vmin(box.mn, t.v3);
        box.mx = vmax(box.mx, t.v1); box.mx = vmax(box.mx, t.v2); box.mx = vmax(box.mx, t.v3);
    }
    m.tree = Tree();
    Node root;
    root.b = box; root.cons = box; root.firstChild = -1; root.refStart = 0; root.refCount = 0;
    m.tree.nodes.push_back(root);
    std::vector<int> all(m.tris.size());
    for (size_t i = 0; i < all.size(); i++) all[i] = (int)i;
    const std::vector<Tri>& tris = m.tris;
    // vertex containment rule (MeshOctree.cs:226-228), inclusive Contains
    auto contains = [&tris](int i, const Box& b) {
        return boxContains(b, tris[i].v1) || boxContains(b, tris[i].v2) || boxContains(b, tris[i].v3);
    };
    return buildRec(m.tree, 0, all, threshold, 0, contains);
}

// OctreeSpatialManager.Build (OctreeSpatialManager.cs:64-99).
static bool buildSceneTree(Ctx& c, int threshold) {
    Box box{V3{0, 0, 0}, V3{0, 0, 0}};
    for (size_t i = 0; i < c.objs.size(); i++) {
        Obj& o = c.objs[i];
        V3 tmn = transform(o.localBox.mn, o.world);
        V3 tmx = transform(o.localBox.mx, o.world);
        Box ob{vmin(tmn, tmx), vmax(tmn, tmx)};
        if (i == 0) box = ob; else box = boxMerge(box, ob);
        // SceneObject.BuildWorld (SceneObject.cs:195-196): Max then Min, no re-sorting
        o.worldBox.mx = transform(o.localBox.mx, o.world);
        o.worldBox.mn = transform(o.localBox.mn, o.world);
    }
    c.sceneTree = Tree();
    Node root;
    root.b = box; root.cons = box; root.firstChild = -1; root.refStart = 0; root.refCount = 0;
    c.sceneTree.nodes.push_back(root);
    std::vector<int> all(c.objs.size());
    for (size_t i = 0; i < all.size(); i++) all[i] = (int)i;
    const std::vector<Obj>& objs = c.objs;
    // WorldBoundingBox.Intersects(child.bounds) (OctreeSpatialManager.cs:240)
    auto contains = [&objs](int i, const Box& b) { return boxIntersects(objs[i].worldBox, b); };
    return buildRec(c.sceneTree, 0, all, threshold, 0, contains);
}

// ------------------------------------------------------------------------------------------
// Triangle test — RayExtensions.IntersectsTriangleBackfaceCulling (RayExtensions.cs:42-75)
// returns 0 = culled by the normal test, 1 = full test failed, 2 = accepted
// ------------------------------------------------------------------------------------------
static inline int triTest(const Ray& ray, const Tri& tri, float& u, float& v, float& dist) {
    u = v = dist = 0;
    V3 D = ray.d;
    V3 T = sub(ray.p, tri.v1);
    float dn = dot(tri.sn, D);
    if (dn > 0) return 0;
    V3 E1 = sub(tri.v2, tri.v1);
    V3 E2 = sub(tri.v3, tri.v1);
    V3 P = cross(D, E2);
    V3 Q = cross(T, E1);
    float row1 = dot(Q, E2);
    float row2 = dot(P, T);
    float row3 = dot(Q, D);
    float inv = 1.0f / dot(P, E1);
    dist = inv * row1;  // (1/det) * Vector3(row1,row2,row3): float * Vector3 is commutative per component
    u = inv * row2;
    v = inv * row3;
    return (u >= 0 && v >= 0 && dist >= 0 && u + v <= 1) ? 2 : 1;
}

// ------------------------------------------------------------------------------------------
// Mesh-level closest hit — MeshOctree.GetRayIntersection (MeshOctree.cs:259-326)
// ------------------------------------------------------------------------------------------
struct MeshHit { int tri; float u, v, d; V3 pos; long long key; };

struct LeafT { float t; int node; };

// GetRayCubeNodeIntersections (MeshOctree.cs:328-353): recursive, all 8 children, no ordering.
static void collectLeaves(const Tree& tree, int n, const Ray& ray, int sem, std::vector<LeafT>& out,
                          LocalCounters& lc) {
    float t;
    lc.nBox++;
    if (!slab(tree.nodes[n].b, ray, t, sem)) return;
    if (tree.nodes[n].firstChild >= 0) {
        for (int i = 0; i < 8; i++) collectLeaves(tree, tree.nodes[n].firstChild + i, ray, sem, out, lc);
    } else {
        out.push_back(LeafT{t, n});
    }
}

// SortedList<float, List<CubeNode>>: ascending keys, leaves with an equal key grouped in
// insertion (DFS) order.  float.CompareTo treats +0 and -0 as equal; keys are never NaN.
static void sortLeaves(std::vector<LeafT>& v) {
    std::stable_sort(v.begin(), v.end(), [](const LeafT& a, const LeafT& b) { return a.t < b.t; });
}

static bool meshTraceLiteral(const Mesh& m, const Ray& ray, int ignoreTri, int sem, MeshHit& out,
                             LocalCounters& lc) {
    std::vector<LeafT> leaves;
    collectLeaves(m.tree, 0, ray, sem, leaves, lc);
    if (leaves.empty()) return false;
    sortLeaves(leaves);
    float minDistance = FLT_MAX, iu = 0, iv = 0;
    int itri = -1;
    bool found = false;
    size_t gi = 0;
    while (!found && gi < leaves.size()) {
        size_t ge = gi + 1;
        while (ge < leaves.size() && leaves[ge].t == leaves[gi].t) ge++;
        for (size_t k = gi; k < ge; k++) {
            const Node& leaf = m.tree.nodes[leaves[k].node];
            for (int i = 0; i < leaf.refCount; i++) {
                int ti = m.tree.refs[leaf.refStart + i];
                lc.nTriRef++;
                if (ignoreTri < 0 || ignoreTri != ti) {
                    float cu, cv, dist;
                    int r = triTest(ray, m.tris[ti], cu, cv, dist);
                    if (r == 0) lc.nTriCulled++; else lc.nTriFull++;
                    if (r == 2 && dist < minDistance) {
                        minDistance = dist; iu = cu; iv = cv; itri = ti; found = true;
                    }
                }
            }
        }
        gi = ge;
    }
    if (found) {
        const Tri& t = m.tris[itri];
        V3 p1 = sub(t.v2, t.v1), p2 = sub(t.v3, t.v1);
        out.tri = itri; out.u = iu; out.v = iv; out.d = minDistance;
        out.pos = add(add(t.v1, mul(p1, iu)), mul(p2, iv));  // MeshOctree.cs:310-312
        out.key = 0;
    }
    return found;
}

// Front-to-back order used by the pruned traversal (and by the CUDA kernels): children are
// visited as (i ^ mask), i = 0..7, mask bit set (4 = x, 2 = y, 1 = z) where the direction
// component is negative.  Any order returns the same answer; the order only fixes the
// counters (N_box_min etc.) so that the GPU and the oracle agree on them exactly.
static inline int octantMask(const Ray& r) {
    return (r.d.x < 0 ? 4 : 0) | (r.d.y < 0 ? 2 : 0) | (r.d.z < 0 ? 1 : 0);
}

// Order-independent restatement: K = min leaf key with an accepted
// triangle; answer = min-d accepted triangle over leaves with key == K, ties -> first in
// (DFS leaf order, list order) == smallest ref position.  Subtrees are pruned when even their
// conservative box (union of descendant boxes: a child's Max can overshoot its parent's by
// rounding, MeshOctree.cs:217-218) is entered after K.
static bool meshTracePruned(const Mesh& m, const Ray& ray, int ignoreTri, int sem, MeshHit& out,
                            LocalCounters& lc) {
    const Tree& tree = m.tree;
    float K = INFINITY;
    bool found = false;
    float bd = 0, bu = 0, bv = 0; int btri = -1; long long bkey = 0;
    struct E { int node; float t; };
    std::vector<E> stack;
    float t0;
    lc.nBox++;
    if (!slab(tree.nodes[0].b, ray, t0, sem)) return false;
    stack.push_back(E{0, t0});
    const int mask = octantMask(ray);
    auto prunable = [&](int n, float t) {
        if (!(t > K)) return false;
        const Node& nd = tree.nodes[n];
        if (nd.firstChild < 0) return true;
        float tc;
        if (!slab(nd.cons, ray, tc, sem)) return false;  // cannot happen (cons ⊇ own); stay safe
        return tc > K;
    };
    while (!stack.empty()) {
        E e = stack.back();
        stack.pop_back();
        if (prunable(e.node, e.t)) continue;
        const Node& nd = tree.nodes[e.node];
        if (nd.firstChild < 0) {
            bool lf = false; float ld = FLT_MAX, lu = 0, lv = 0; int lt = -1; long long lk = 0;
            for (int i = 0; i < nd.refCount; i++) {
                int ti = tree.refs[nd.refStart + i];
                lc.nTriRef++;
                if (ignoreTri < 0 || ignoreTri != ti) {
                    float cu, cv, dist;
                    int r = triTest(ray, m.tris[ti], cu, cv, dist);
                    if (r == 0) lc.nTriCulled++; else lc.nTriFull++;
                    if (r == 2 && dist < ld) { ld = dist; lu = cu; lv = cv; lt = ti; lk = nd.refStart + i; lf = true; }
                }
            }
            if (lf) {
                if (!found || e.t < K) { found = true; K = e.t; bd = ld; bu = lu; bv = lv; btri = lt; bkey = lk; }
                else if (e.t == K && (ld < bd || (ld == bd && lk < bkey))) { bd = ld; bu = lu; bv = lv; btri = lt; bkey = lk; }
            }
        } else {
            for (int i = 7; i >= 0; i--) {  // push farthest first so the nearest pops first
                int c = nd.firstChild + (i ^ mask);
                float tc;
                lc.nBox++;
                if (slab(tree.nodes[c].b, ray, tc, sem) && !prunable(c, tc)) stack.push_back(E{c, tc});
            }
        }
    }
    if (found) {
        const Tri& t = m.tris[btri];
        V3 p1 = sub(t.v2, t.v1), p2 = sub(t.v3, t.v1);
        out.tri = btri; out.u = bu; out.v = bv; out.d = bd;
        out.pos = add(add(t.v1, mul(p1, bu)), mul(p2, bv));
        out.key = bkey;
    }
    return found;
}

// ------------------------------------------------------------------------------------------
// Scene-level closest hit — OctreeSpatialManager.GetRayIntersection (OctreeSpatialManager.cs:312-455)
// ignoreMesh/ignoreTri identify the Triangle object passed as ignoreTriangle: Triangle
// instances are owned by a Mesh that all instances of a model share (SceneObject.cs:126-127),
// so identity == (mesh id, index in Mesh.Triangles) — NOT the instance (Q6).  The reference's
// ignoreObject parameter is dead (Q5).
// ------------------------------------------------------------------------------------------
struct ObjHit { bool found; int mesh; MeshHit mh; };

// Body of the per-object block (OctreeSpatialManager.cs:347-428).
static inline void objectQuery(const Ctx& c, int objIdx, const Ray& ray, int ignoreMesh, int ignoreTri,
                               float& minDistance, bool& found, Hit& best, V3& bestObjPos,
                               long long& bestKey, long long keyBase, LocalCounters& lc) {
    const Obj& o = c.objs[objIdx];
    V3 rdp = add(ray.p, ray.d);
    V3 v1 = transform(ray.p, o.inv);
    V3 v2 = transform(rdp, o.inv);
    Ray tr{v1, normalize(sub(v2, v1))};
    if (c.mode == MODE_PRUNED && (!std::isfinite(tr.p.x) || !std::isfinite(tr.p.y) || !std::isfinite(tr.p.z) ||
                                  !std::isfinite(tr.d.x) || !std::isfinite(tr.d.y) || !std::isfinite(tr.d.z)))
        return;  // same shortcut one level down (object-space ray)
    for (size_t mi = 0; mi < o.meshes.size(); mi++) {
        const Mesh& m = *c.meshes[o.meshes[mi]];
        float tb;
        lc.nBox++;
        if (!slab(m.box, tr, tb, c.semantics)) continue;  // Mesh.RayIntersects (Mesh.cs:34-39)
        MeshHit mh{};
        int ign = (ignoreMesh == o.meshes[mi]) ? ignoreTri : -1;
        bool ok = (c.mode == MODE_PRUNED) ? meshTracePruned(m, tr, ign, c.semantics, mh, lc)
                                          : meshTraceLiteral(m, tr, ign, c.semantics, mh, lc);
        long long key = keyBase + (long long)mi;
        if (ok && (mh.d < minDistance || (c.mode == MODE_PRUNED && found && mh.d == minDistance && key < bestKey))) {
            minDistance = mh.d;
            best.obj = objIdx; best.mesh = o.meshes[mi]; best.tri = mh.tri;
            best.u = mh.u; best.v = mh.v; best.d = mh.d;
            bestObjPos = mh.pos;
            bestKey = key;
            found = true;
        }
    }
}

static bool sceneTraceLiteral(const Ctx& c, const Ray& ray, int ignoreMesh, int ignoreTri, Hit& out,
                              LocalCounters& lc) {
    std::vector<LeafT> leaves;
    collectLeaves(c.sceneTree, 0, ray, c.semantics, leaves, lc);
    if (leaves.empty()) return false;
    sortLeaves(leaves);
    float minDistance = FLT_MAX;
    bool found = false;
    Hit best; V3 bestPos{0, 0, 0}; long long bestKey = 0;
    size_t gi = 0;
    while (!found && gi < leaves.size()) {
        size_t ge = gi + 1;
        while (ge < leaves.size() && leaves[ge].t == leaves[gi].t) ge++;
        for (size_t k = gi; k < ge; k++) {
            const Node& leaf = c.sceneTree.nodes[leaves[k].node];
            for (int i = 0; i < leaf.refCount; i++)
                objectQuery(c, c.sceneTree.refs[leaf.refStart + i], ray, ignoreMesh, ignoreTri, minDistance,
                            found, best, bestPos, bestKey, 0, lc);
        }
        gi = ge;
    }
    if (found) {
        out = best;
        out.worldPos = transform(bestPos, c.objs[best.obj].world);  // OctreeSpatialManager.cs:441-443
    }
    return found;
}

static inline bool finite3(V3 v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); }

static bool sceneTracePruned(const Ctx& c, const Ray& ray, int ignoreMesh, int ignoreTri, Hit& out,
                             LocalCounters& lc) {
    // A ray with a NaN/inf component can never pass the triangle test (every comparison on the NaN
    // products is false), so the reference walks the whole tree and reports a miss; the pruned
    // restatement (and the CUDA kernels) report the miss without doing that work.
    if (!finite3(ray.p) || !finite3(ray.d)) return false;
    const Tree& tree = c.sceneTree;
    float K = INFINITY;
    bool found = false;
    Hit best; V3 bestPos{0, 0, 0}; long long bestKey = 0; float bestD = 0;
    struct E { int node; float t; };
    std::vector<E> stack;
    float t0;
    lc.nBox++;
    if (!slab(tree.nodes[0].b, ray, t0, c.semantics)) return false;
    stack.push_back(E{0, t0});
    const int mask = octantMask(ray);
    auto prunable = [&](int n, float t) {
        if (!(t > K)) return false;
        const Node& nd = tree.nodes[n];
        if (nd.firstChild < 0) return true;
        float tc;
        if (!slab(nd.cons, ray, tc, c.semantics)) return false;
        return tc > K;
    };
    while (!stack.empty()) {
        E e = stack.back();
        stack.pop_back();
        if (prunable(e.node, e.t)) continue;
        const Node& nd = tree.nodes[e.node];
        if (nd.firstChild < 0) {
            float ld = FLT_MAX; bool lf = false; Hit lh; V3 lpos{0, 0, 0}; long long lkey = 0;
            for (int i = 0; i < nd.refCount; i++)
                objectQuery(c, tree.refs[nd.refStart + i], ray, ignoreMesh, ignoreTri, ld, lf, lh, lpos, lkey,
                            ((long long)(nd.refStart + i)) << 20, lc);
            if (lf) {
                if (!found || e.t < K) { found = true; K = e.t; best = lh; bestPos = lpos; bestKey = lkey; bestD = ld; }
                else if (e.t == K && (ld < bestD || (ld == bestD && lkey < bestKey))) { best = lh; bestPos = lpos; bestKey = lkey; bestD = ld; }
            }
        } else {
            for (int i = 7; i >= 0; i--) {
                int ch = nd.firstChild + (i ^ mask);
                float tc;
                lc.nBox++;
                if (slab(tree.nodes[ch].b, ray, tc, c.semantics) && !prunable(ch, tc)) stack.push_back(E{ch, tc});
            }
        }
    }
    if (found) {
        out = best;
        out.worldPos = transform(bestPos, c.objs[best.obj].world);
    }
    return found;
}

static inline bool sceneTrace(const Ctx& c, const Ray& ray, int ignoreMesh, int ignoreTri, Hit& out,
                              LocalCounters& lc) {
    return c.mode == MODE_PRUNED ? sceneTracePruned(c, ray, ignoreMesh, ignoreTri, out, lc)
                                 : sceneTraceLiteral(c, ray, ignoreMesh, ignoreTri, out, lc);
}

// ------------------------------------------------------------------------------------------
// Texture lookup — Material.LookupUV (Material.cs:71-232)
// ------------------------------------------------------------------------------------------
static const float BYTE_RECIPROCAL = 1.0f / 255.0f;  // Material.cs:28

static inline void wrapUV(V2& uv) {  // Material.cs:125-136
    if (uv.x > 1.0f) uv.x = std::fmod(uv.x, 1.0f);
    if (uv.y > 1.0f) uv.y = std::fmod(uv.y, 1.0f);
    if (uv.x < 0.0f) uv.x = 1.0f + std::fmod(uv.x, 1.0f);
    if (uv.y < 0.0f) uv.y = 1.0f + std::fmod(uv.y, 1.0f);
}
static inline void mirrorUV(V2& uv) {  // Material.cs:102-123
    V2 orig = uv;
    wrapUV(uv);
    if ((int)(orig.x - uv.x) % 2 == 0) uv.x = 1.0f - uv.x;
    if ((int)(orig.y - uv.y) % 2 == 0) uv.y = 1.0f - uv.y;
}
static inline void clampUV(V2& uv) {  // Vector2.Clamp(uv, 0, 1): min(max(v, lo), hi) with XNA compare form
    float x = uv.x; x = (x > 1.0f) ? 1.0f : x; x = (x < 0.0f) ? 0.0f : x; uv.x = x;
    float y = uv.y; y = (y > 1.0f) ? 1.0f : y; y = (y < 0.0f) ? 0.0f : y; uv.y = y;
}
static inline int toIntTrunc(float f) {  // C# (int)float; out-of-range/NaN -> 0x80000000 on x86/x64
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT32_MIN;
    return (int)f;
}
static inline uint32_t texel(const Material& m, int x, int y) {
    // The reference reads through an unsafe pointer (Material.cs:150-152) and would fault on an
    // out-of-range index; finite UVs never produce one.  Clamp so the oracle cannot crash.
    if (x < 0) x = 0; if (x >= m.texW) x = m.texW - 1;
    if (y < 0) y = 0; if (y >= m.texH) y = m.texH - 1;
    return m.argb[(size_t)m.texW * (size_t)y + (size_t)x];
}
static V3 lookupUV(const Material& m, V2 uv, int addressMode, int filtering) {
    switch (addressMode) {
        case XNART_ADDRESS_CLAMP: clampUV(uv); break;
        case XNART_ADDRESS_WRAP: wrapUV(uv); break;
        default: mirrorUV(uv); break;
    }
    if (filtering == XNART_FILTER_POINT) {  // GetColorPoint, Material.cs:145-160
        int x = toIntTrunc(uv.x * (float)(m.texW - 1));
        int y = toIntTrunc(uv.y * (float)(m.texH - 1));
        uint32_t a = texel(m, x, y);
        return V3{(float)((a >> 16) & 255u) * BYTE_RECIPROCAL, (float)((a >> 8) & 255u) * BYTE_RECIPROCAL,
                  (float)(a & 255u) * BYTE_RECIPROCAL};
    }
    // GetColorBilinear, Material.cs:162-232
    float tdx = 1.0f / (float)m.texW, tdy = 1.0f / (float)m.texH;  // texelDensity, Material.cs:67
    double remX = std::remainder((double)uv.x, (double)tdx);
    double remY = std::remainder((double)uv.y, (double)tdy);
    uv.x -= (float)remX;
    uv.y -= (float)remY;
    int x = toIntTrunc(uv.x * (float)(m.texW - 1));
    int y = toIntTrunc(uv.y * (float)(m.texH - 1));
    int x2 = toIntTrunc((uv.x + tdx) * (float)(m.texW - 1));
    int y2 = toIntTrunc((uv.y + tdy) * (float)(m.texH - 1));
    auto rgb = [](uint32_t a) { return V3{(float)((a >> 16) & 255u), (float)((a >> 8) & 255u), (float)(a & 255u)}; };
    V3 base = rgb(texel(m, x, y)), bx = rgb(texel(m, x2, y)), by = rgb(texel(m, x, y2)), bxy = rgb(texel(m, x2, y2));
    float dx = (float)(remX * (double)m.texW) + 0.5f;
    float dy = (float)(remY * (double)m.texH) + 0.5f;
    float idx = 1.0f - dx, idy = 1.0f - dy;
    V3 s = add(add(add(mul(mul(base, idx), idy), mul(mul(by, idx), dy)), mul(mul(bx, dx), idy)), mul(mul(bxy, dx), dy));
    return mul(s, BYTE_RECIPROCAL);
}

// ------------------------------------------------------------------------------------------
// Lights — SpotLight.GetLightForFragment (SpotLight.cs:37-62), DirectionalLight (:23-30)
// ------------------------------------------------------------------------------------------
static V3 lightForFragment(const Light& L, V3 position, V3 normal) {
    if (L.kind == XNART_LIGHT_DIRECTIONAL) {
        float sd = dot(L.direction, normal);
        if (sd < 0.0f) sd = 0.0f;
        return mul(mul(L.color, sd), L.intensity);
    }
    V3 toL = normalize(sub(L.position, position));
    float sd = dot(toL, normal);
    if (sd < 0.0f) return V3{0, 0, 0};
    float ld = dot(neg(toL), L.direction);
    if (ld > L.angleCosine) {
        float spot = L.intensity * (float)((double)(ld - L.angleCosine) /
                                           std::pow((double)(1.0f - L.angleCosine), (double)L.decayExponent));
        float hl = (float)std::pow((double)sd, 12.0);
        V3 a = mul(mul(L.color, spot), sd);
        return add(a, V3{1.0f * hl, 1.0f * hl, 1.0f * hl});
    }
    return V3{0, 0, 0};
}

// ------------------------------------------------------------------------------------------
// CastRay — RayTracer.cs:506-737 (+ IsLightPathObstructed :465-502)
// ------------------------------------------------------------------------------------------
struct FrameParams {
    int maxReflections, filtering, addressMode, semantics;
};

static float lightPathObstructed(const Ctx& c, const Hit& result, const Light& L, LocalCounters& lc) {
    V3 dirToLight; float dist;
    if (L.kind == XNART_LIGHT_SPOT) {
        dirToLight = sub(L.position, result.worldPos);
        dist = length(dirToLight);
        dirToLight = normalize(dirToLight);
    } else {
        dirToLight = neg(L.direction);
        dist = FLT_MAX;
    }
    Ray shadow{result.worldPos, dirToLight};
    Hit sh;
    lc.raysShadow++;
    if (sceneTrace(c, shadow, result.mesh, result.tri, sh, lc) && sh.d < dist) {
        const Mesh& m = *c.meshes[sh.mesh];
        return m.mat.transparent ? m.tris[sh.tri].color.w : 1.0f;
    }
    return 0.0f;
}

static V3 surfaceColorOf(const Mesh& m, const Tri& t, const Hit& r, const FrameParams& fp) {
    if (m.mat.useTexture) {
        V2 a{t.uv2.x - t.uv1.x, t.uv2.y - t.uv1.y};
        V2 b{t.uv3.x - t.uv1.x, t.uv3.y - t.uv1.y};
        V2 uv{(t.uv1.x + (a.x * r.u)) + (b.x * r.v), (t.uv1.y + (a.y * r.u)) + (b.y * r.v)};
        return lookupUV(m.mat, uv, fp.addressMode, fp.filtering);
    }
    return V3{t.color.x, t.color.y, t.color.z};
}

static uint32_t castRay(const Ctx& c, const FrameParams& fp, const Ray& ray, int iteration, int originMesh,
                        int originTri, float currentRefIndex, LocalCounters& lc) {
    Hit result;
    if (!sceneTrace(c, ray, originMesh, originTri, result, lc))
        return packColor(V3{0, 0, 0}, fp.semantics);  // RayTracer.cs:732
    lc.hits++;
    const Mesh& mesh = *c.meshes[result.mesh];
    const Material& mat = mesh.mat;
    const Tri& tri = mesh.tris[result.tri];
    V3 N;
    if (mat.interpolateNormals) {
        V3 n1 = sub(tri.n2, tri.n1), n2 = sub(tri.n3, tri.n1);
        N = normalize(add(add(tri.n1, mul(n1, result.u)), mul(n2, result.v)));
    } else {
        N = tri.sn;
    }
    V3 lightResult{0, 0, 0};
    for (const Light& L : c.lights) {
        float amount = lightPathObstructed(c, result, L, lc);
        if (amount != 1.0f) lightResult = add(lightResult, mul(lightForFragment(L, result.worldPos, N), 1.0f - amount));
    }
    if (iteration < fp.maxReflections) {
        Ray r;
        r.p = result.worldPos;
        r.d = normalize(reflect(ray.d, N));
        lc.raysReflection++;
        // convexGeometry is never set anywhere in the reference -> always the `null` ignoreObject branch
        uint32_t reflColor = castRay(c, fp, r, iteration + 1, result.mesh, result.tri, currentRefIndex, lc);
        V3 surf = surfaceColorOf(mesh, tri, result, fp);
        V3 cv = mul(lerp(unpackColor(reflColor), surf, 1.0f - mat.reflectiveness), lightResult);
        if (mat.transparent) {
            float n1, n2;
            if (currentRefIndex == mat.refractionIndex) { n1 = 1.0f; n2 = currentRefIndex; }
            else { n1 = mat.refractionIndex; n2 = 1.0f; }
            float cos1 = dot(N, neg(ray.d));
            double ratio = (double)(n1 / n2);
            double c1 = (double)cos1;
            float cos2 = (float)std::sqrt(1.0 - (ratio * ratio) * (1.0 - (c1 * c1)));  // Math.Pow(x, 2.0) == x*x exactly for float-valued x
            float k = n1 / n2;
            V3 refr;
            if (cos1 >= 0) refr = add(mul(ray.d, k), mul(N, (k * cos1) - cos2));
            else refr = sub(mul(ray.d, k), mul(N, (k * cos1) - cos2));
            Ray rr{result.worldPos, normalize(refr)};
            lc.raysRefraction++;
            uint32_t refrColor = castRay(c, fp, rr, iteration + 1, result.mesh, result.tri, n2, lc);
            cv = lerp(unpackColor(refrColor), cv, tri.color.w);
        }
        return packColor(cv, fp.semantics);
    }
    V3 surf = surfaceColorOf(mesh, tri, result, fp);
    return packColor(mul(lightResult, surf), fp.semantics);
}

// ------------------------------------------------------------------------------------------
// Primary rays — RayTracer.Render (RayTracer.cs:391-429)
// ------------------------------------------------------------------------------------------
static M4 toM4(const float* p) { M4 m; std::memcpy(m.m, p, 64); return m; }

static M4 frameInverse(const XnartFrame& f) {
    // Viewport.Unproject: Invert(Multiply(Multiply(world = Identity, view), proj))
    return invert(multiply(multiply(identity(), toM4(f.view)), toM4(f.proj)));
}

static Ray primaryRay(const XnartFrame& f, const M4& inv, int x, int y) {
    Viewport vp{f.vpX, f.vpY, f.vpW, f.vpH, f.minDepth, f.maxDepth};
    Ray r;
    r.p = unprojectWithInverse(vp, V3{(float)x, (float)y, 0.0f}, inv);
    V3 farP = unprojectWithInverse(vp, V3{(float)x, (float)y, 1.0f}, inv);
    r.d = normalize(sub(farP, r.p));
    return r;
}

static void flush(Counters& g, const LocalCounters& l) {
    g.raysPrimary += l.raysPrimary; g.raysShadow += l.raysShadow; g.raysReflection += l.raysReflection;
    g.raysRefraction += l.raysRefraction; g.hits += l.hits; g.nBox += l.nBox; g.nTriRef += l.nTriRef;
    g.nTriCulled += l.nTriCulled; g.nTriFull += l.nTriFull; g.pixels += l.pixels;
}

}  // namespace xo

// ==========================================================================================
// C API (mirrors include/xnart.h with the xo_ prefix; same struct layouts)
// ==========================================================================================
using namespace xo;

extern "C" {

struct XoCtx { Ctx c; int slabOverride = -1; };

int32_t xo_create(XoCtx** out) { *out = new XoCtx(); return 0; }
int32_t xo_destroy(XoCtx* h) { delete h; return 0; }
const char* xo_last_error(XoCtx* h) { return h->c.err.c_str(); }
// mode: 0 literal (reference algorithm), 1 pruned (order-independent restatement)
int32_t xo_set_mode(XoCtx* h, int32_t mode) { h->c.mode = mode; return 0; }
int32_t xo_set_semantics(XoCtx* h, int32_t s) { h->c.semantics = s; return 0; }
// Slab-test form used while rendering: -1 = follow XnartFrame.semantics (a pure XNA 4.0 or pure MonoGame framework),
// 0 / 1 = pinned.  libxnart's device code always runs the XNA 4.0 form, whatever colour pack a frame selects.
int32_t xo_set_slab_override(XoCtx* h, int32_t s) { h->slabOverride = s; return 0; }

static V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }

int32_t xo_add_mesh(XoCtx* h, const XnartTriangle* tris, int64_t nTri, const XnartMaterial* material,
                    const float box[6], int32_t itemThreshold, int32_t* meshIdOut) {
    auto m = std::make_unique<Mesh>();
    m->tris.resize((size_t)nTri);
    for (int64_t i = 0; i < nTri; i++) {
        const XnartTriangle& s = tris[i];
        Tri& t = m->tris[(size_t)i];
        t.v1 = v3(s.v1); t.v2 = v3(s.v2); t.v3 = v3(s.v3);
        t.uv1 = V2{s.uv1[0], s.uv1[1]}; t.uv2 = V2{s.uv2[0], s.uv2[1]}; t.uv3 = V2{s.uv3[0], s.uv3[1]};
        t.n1 = v3(s.n1); t.n2 = v3(s.n2); t.n3 = v3(s.n3);
        t.sn = v3(s.surfaceNormal);
        t.color = V4{s.color[0], s.color[1], s.color[2], s.color[3]};
        t.id = s.id;
    }
    m->mat.reflectiveness = material->reflectiveness;
    m->mat.useTexture = material->useTexture != 0;
    m->mat.interpolateNormals = material->interpolateNormals != 0;
    m->mat.transparent = material->transparent != 0;
    m->mat.refractionIndex = material->refractionIndex;
    m->mat.texW = material->texW; m->mat.texH = material->texH;
    if (m->mat.useTexture) {
        if (!material->argb || material->texW <= 0 || material->texH <= 0) { h->c.err = "useTexture without texels"; return XNART_ERR_INVALID_ARGUMENT; }
        m->mat.argb.assign(material->argb, material->argb + (size_t)material->texW * (size_t)material->texH);
    }
    m->box = Box{v3(box), v3(box + 3)};
    if (!buildMeshTree(*m, itemThreshold)) { h->c.err = "mesh octree exceeds depth guard"; return XNART_ERR_OCTREE_DEPTH; }
    *meshIdOut = (int32_t)h->c.meshes.size();
    h->c.meshes.push_back(std::move(m));
    h->c.built = false;
    return 0;
}

int32_t xo_add_object(XoCtx* h, const int32_t* meshIds, int32_t nMesh, const float world[16],
                      const float inverseWorld[16], const float localBox[6], int32_t* objIdOut) {
    Obj o;
    for (int i = 0; i < nMesh; i++) {
        if (meshIds[i] < 0 || meshIds[i] >= (int)h->c.meshes.size()) { h->c.err = "bad mesh id"; return XNART_ERR_INVALID_ARGUMENT; }
        o.meshes.push_back(meshIds[i]);
    }
    o.world = toM4(world); o.inv = toM4(inverseWorld);
    o.localBox = Box{v3(localBox), v3(localBox + 3)};
    o.worldBox = o.localBox;
    *objIdOut = (int32_t)h->c.objs.size();
    h->c.objs.push_back(o);
    h->c.built = false;
    return 0;
}

// scene.Bodies.Clear() / a body moved after Build(): the reference reads SceneObject.World / InverseWorld per ray
// (OctreeSpatialManager.cs:358-364, 441-443) and never rebuilds its tree on its own.
int32_t xo_clear_objects(XoCtx* h) {
    h->c.objs.clear();
    h->c.built = false;
    return 0;
}
int32_t xo_set_object_transform(XoCtx* h, int32_t objId, const float world[16], const float inverseWorld[16]) {
    if (objId < 0 || objId >= (int)h->c.objs.size()) { h->c.err = "bad object id"; return XNART_ERR_INVALID_ARGUMENT; }
    h->c.objs[(size_t)objId].world = toM4(world);
    h->c.objs[(size_t)objId].inv = toM4(inverseWorld);
    return 0;
}

int32_t xo_build_scene(XoCtx* h, int32_t itemThreshold) {
    if (!buildSceneTree(h->c, itemThreshold)) { h->c.err = "scene octree exceeds depth guard"; return XNART_ERR_OCTREE_DEPTH; }
    h->c.built = true;
    return 0;
}

int32_t xo_set_lights(XoCtx* h, const XnartLight* lights, int32_t n) {
    h->c.lights.clear();
    for (int i = 0; i < n; i++) {
        Light L;
        L.kind = lights[i].kind;
        L.position = v3(lights[i].position); L.direction = v3(lights[i].direction); L.color = v3(lights[i].color);
        L.intensity = lights[i].intensity; L.decayExponent = lights[i].decayExponent; L.spotAngle = lights[i].spotAngle;
        L.angleCosine = (float)std::cos((double)(L.spotAngle * 0.5f));  // SpotLight.cs:25
        h->c.lights.push_back(L);
    }
    return 0;
}

static void resetCounters(Counters& g) {
    g.raysPrimary = 0; g.raysShadow = 0; g.raysReflection = 0; g.raysRefraction = 0; g.hits = 0;
    g.nBox = 0; g.nTriRef = 0; g.nTriCulled = 0; g.nTriFull = 0; g.pixels = 0;
}

// RenderInternal + Render (RayTracer.cs:103-126, 391-429).  nThreads workers share the
// reference's scanline dispenser (Interlocked.Increment, RayTracer.cs:49-52); the reference
// as shipped uses 1 (RayTracer.cs:99).  y0/y1 restrict the scanline range (bounded samples
// for the CPU baseline); pass 0, vpH for a full frame.
int32_t xo_render_rows(XoCtx* h, const XnartFrame* f, uint32_t* out, int32_t nThreads, int32_t y0, int32_t y1) {
    Ctx& c = h->c;
    if (!c.built) { c.err = "scene not built"; return XNART_ERR_NOT_BUILT; }
    if (f->filtering < 0 || f->filtering > 1 || f->addressMode < 0 || f->addressMode > 2) { c.err = "bad enum"; return XNART_ERR_INVALID_ARGUMENT; }
    resetCounters(c.counters);
    const int savedSem = c.semantics;
    c.semantics = h->slabOverride >= 0 ? h->slabOverride : f->semantics;
    FrameParams fp{f->maxReflections, f->filtering, f->addressMode, f->semantics};
    const M4 inv = frameInverse(*f);
    std::atomic<int> scanline{y0 - 1};
    const int W = f->vpW;
    auto worker = [&]() {
        LocalCounters lc;
        for (;;) {
            int y = ++scanline;
            if (y >= y1) break;
            if (f->shardCount > 1) {
                int band = y / (f->shardBandRows > 0 ? f->shardBandRows : 1);
                if (band % f->shardCount != f->shardIndex) continue;
            }
            for (int x = 0; x < W; x++) {
                Ray r = primaryRay(*f, inv, x, y);
                lc.raysPrimary++; lc.pixels++;
                out[(size_t)y * (size_t)W + (size_t)x] = castRay(c, fp, r, 0, -1, -1, 1.0f, lc);
            }
        }
        flush(c.counters, lc);
    };
    if (nThreads <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nThreads; i++) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    c.semantics = savedSem;
    return 0;
}

int32_t xo_render_frame(XoCtx* h, const XnartFrame* f, uint32_t* out, int32_t nThreads) {
    return xo_render_rows(h, f, out, nThreads, 0, f->vpH);
}

// ------------------------------------------------------------------------------------------
// Adaptive multisampling — RayTracer.RenderInternalWithMultisampling / RenderFirstPass /
// GetColorForQuadrant (RayTracer.cs:128-311), literal restatement including the :305 slip (the
// refined lower-right quadrant is stored into urColor, lrColor keeps its sample).
// ------------------------------------------------------------------------------------------
static Ray sampleRay(const XnartFrame& f, const M4& inv, float sx, float sy) {
    Viewport vp{f.vpX, f.vpY, f.vpW, f.vpH, f.minDepth, f.maxDepth};
    Ray r;
    r.p = unprojectWithInverse(vp, V3{sx, sy, 0.0f}, inv);     // :224-227
    V3 v2 = unprojectWithInverse(vp, V3{sx, sy, 1.0f}, inv);    // :229-230
    r.d = normalize(sub(v2, r.p));                              // :231-232
    return r;
}

static uint32_t colorForQuadrant(const Ctx& c, const FrameParams& fp, const XnartFrame& f, const M4& inv, float centerX, float centerY,
                                 float size, int iteration, int quality, LocalCounters& lc) {
    const float TRESHOLD = 0.5f;                                  // :340
    const float quarterSize = size * 0.25f;                       // :218
    const float xs[4] = {centerX - quarterSize, centerX + quarterSize, centerX - quarterSize, centerX + quarterSize};
    const float ys[4] = {centerY - quarterSize, centerY - quarterSize, centerY + quarterSize, centerY + quarterSize};
    uint32_t col[4];                                              // ul, ur, ll, lr
    for (int k = 0; k < 4; k++) {
        Ray r = sampleRay(f, inv, xs[k], ys[k]);
        lc.raysPrimary++;
        col[k] = castRay(c, fp, r, 0, -1, -1, 1.0f, lc);          // :234, 248, 262, 276
    }
    if (iteration < quality) {                                    // :279
        V3 v[4];
        for (int k = 0; k < 4; k++) v[k] = unpackColor(col[k]);
        V3 average = divs(add(add(add(v[0], v[1]), v[2]), v[3]), 4.0f);   // :285
        float averageLength = length(average);
        const float half = size / 2.0f;
        if (std::fabs(averageLength - length(v[0])) > TRESHOLD)
            col[0] = colorForQuadrant(c, fp, f, inv, centerX - quarterSize, centerY - quarterSize, half, iteration + 1, quality, lc);
        if (std::fabs(averageLength - length(v[1])) > TRESHOLD)
            col[1] = colorForQuadrant(c, fp, f, inv, centerX + quarterSize, centerY - quarterSize, half, iteration + 1, quality, lc);
        if (std::fabs(averageLength - length(v[2])) > TRESHOLD)
            col[2] = colorForQuadrant(c, fp, f, inv, centerX - quarterSize, centerY + quarterSize, half, iteration + 1, quality, lc);
        if (std::fabs(averageLength - length(v[3])) > TRESHOLD)   // :303-306: the result lands in urColor
            col[1] = colorForQuadrant(c, fp, f, inv, centerX + quarterSize, centerY + quarterSize, half, iteration + 1, quality, lc);
    }
    V3 sum = add(add(add(unpackColor(col[0]), unpackColor(col[1])), unpackColor(col[2])), unpackColor(col[3]));
    return packColor(divs(sum, 4.0f), fp.semantics);              // :309
}

// CastRay for arbitrary screen-space sample positions (the four casts of GetColorForQuadrant, :220-276).
int32_t xo_cast_samples(XoCtx* h, const XnartFrame* f, const float* xy, int64_t n, uint32_t* out) {
    Ctx& c = h->c;
    if (!c.built) { c.err = "scene not built"; return XNART_ERR_NOT_BUILT; }
    const int savedSem = c.semantics;
    c.semantics = h->slabOverride >= 0 ? h->slabOverride : f->semantics;
    FrameParams fp{f->maxReflections, f->filtering, f->addressMode, f->semantics};
    const M4 inv = frameInverse(*f);
    LocalCounters lc;
    for (int64_t i = 0; i < n; i++) {
        Ray r = sampleRay(*f, inv, xy[2 * i], xy[2 * i + 1]);
        out[i] = castRay(c, fp, r, 0, -1, -1, 1.0f, lc);
    }
    c.semantics = savedSem;
    return 0;
}

int32_t xo_render_frame_multisample(XoCtx* h, const XnartFrame* f, int32_t quality, uint32_t* out, int32_t nThreads) {
    Ctx& c = h->c;
    if (!c.built) { c.err = "scene not built"; return XNART_ERR_NOT_BUILT; }
    resetCounters(c.counters);
    const int savedSem = c.semantics;
    c.semantics = h->slabOverride >= 0 ? h->slabOverride : f->semantics;
    FrameParams fp{f->maxReflections, f->filtering, f->addressMode, f->semantics};
    const M4 inv = frameInverse(*f);
    std::atomic<int> scanline{-1};
    const int W = f->vpW, H = f->vpH;
    auto worker = [&]() {
        LocalCounters lc;
        for (;;) {
            int y = ++scanline;
            if (y >= H) break;
            for (int x = 0; x < W; x++) {
                lc.pixels++;
                out[(size_t)y * (size_t)W + (size_t)x] = colorForQuadrant(c, fp, *f, inv, (float)x, (float)y, 1.0f, 0, quality, lc);  // :197
            }
        }
        flush(c.counters, lc);
    };
    if (nThreads <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nThreads; i++) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    c.semantics = savedSem;
    return 0;
}

int32_t xo_primary_rays(const XnartFrame* f, float* rays6) {
    const M4 inv = frameInverse(*f);
    for (int y = 0; y < f->vpH; y++)
        for (int x = 0; x < f->vpW; x++) {
            Ray r = primaryRay(*f, inv, x, y);
            float* o = rays6 + 6 * ((size_t)y * (size_t)f->vpW + (size_t)x);
            o[0] = r.p.x; o[1] = r.p.y; o[2] = r.p.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
        }
    return 0;
}

int32_t xo_trace_rays(XoCtx* h, const float* rays6, int64_t n, const int32_t* ignore2, int32_t* objId,
                      int32_t* meshId, int32_t* triId, float* u, float* v, float* d, float* worldPos3) {
    Ctx& c = h->c;
    if (!c.built) { c.err = "scene not built"; return XNART_ERR_NOT_BUILT; }
    resetCounters(c.counters);
    LocalCounters lc;
    for (int64_t i = 0; i < n; i++) {
        Ray r{v3(rays6 + 6 * i), v3(rays6 + 6 * i + 3)};
        int im = ignore2 ? ignore2[2 * i] : -1, it = ignore2 ? ignore2[2 * i + 1] : -1;
        Hit hit;
        lc.raysPrimary++;
        bool ok = sceneTrace(c, r, im, it, hit, lc);
        if (ok) lc.hits++;
        objId[i] = ok ? hit.obj : -1; meshId[i] = ok ? hit.mesh : -1; triId[i] = ok ? hit.tri : -1;
        u[i] = ok ? hit.u : 0; v[i] = ok ? hit.v : 0; d[i] = ok ? hit.d : 0;
        if (worldPos3) { worldPos3[3 * i] = ok ? hit.worldPos.x : 0; worldPos3[3 * i + 1] = ok ? hit.worldPos.y : 0; worldPos3[3 * i + 2] = ok ? hit.worldPos.z : 0; }
    }
    flush(c.counters, lc);
    return 0;
}

int32_t xo_get_counters(XoCtx* h, XnartCounters* out) {
    std::memset(out, 0, sizeof(*out));
    const Counters& g = h->c.counters;
    out->raysPrimary = g.raysPrimary; out->raysShadow = g.raysShadow; out->raysReflection = g.raysReflection;
    out->raysRefraction = g.raysRefraction; out->hits = g.hits; out->nBox = g.nBox; out->nTriRef = g.nTriRef;
    out->nTriCulled = g.nTriCulled; out->nTriFull = g.nTriFull; out->pixels = g.pixels;
    return 0;
}

// ---- introspection / unit helpers for the known-answer tests ----------------------------------
// which: 0 = mesh tree `id`, 1 = scene tree
static const Tree* pickTree(XoCtx* h, int which, int id) {
    if (which == 1) return &h->c.sceneTree;
    if (id < 0 || id >= (int)h->c.meshes.size()) return nullptr;
    return &h->c.meshes[id]->tree;
}
int32_t xo_tree_info(XoCtx* h, int32_t which, int32_t id, int32_t* nNodes, int32_t* nRefs, int32_t* depth) {
    const Tree* t = pickTree(h, which, id);
    if (!t) return XNART_ERR_INVALID_ARGUMENT;
    *nNodes = (int32_t)t->nodes.size(); *nRefs = (int32_t)t->refs.size(); *depth = t->depth;
    return 0;
}
int32_t xo_tree_dump(XoCtx* h, int32_t which, int32_t id, float* boxes6, float* cons6, int32_t* firstChild,
                     int32_t* refStart, int32_t* refCount, int32_t* refs) {
    const Tree* t = pickTree(h, which, id);
    if (!t) return XNART_ERR_INVALID_ARGUMENT;
    for (size_t i = 0; i < t->nodes.size(); i++) {
        const Node& n = t->nodes[i];
        const float b[6] = {n.b.mn.x, n.b.mn.y, n.b.mn.z, n.b.mx.x, n.b.mx.y, n.b.mx.z};
        const float cb[6] = {n.cons.mn.x, n.cons.mn.y, n.cons.mn.z, n.cons.mx.x, n.cons.mx.y, n.cons.mx.z};
        std::memcpy(boxes6 + 6 * i, b, 24);
        if (cons6) std::memcpy(cons6 + 6 * i, cb, 24);
        firstChild[i] = n.firstChild; refStart[i] = n.refStart; refCount[i] = n.refCount;
    }
    if (refs) std::memcpy(refs, t->refs.data(), t->refs.size() * sizeof(int));
    return 0;
}

int32_t xo_slab(const float box6[6], const float ray6[6], int32_t semantics, float* t) {
    Box b{v3(box6), v3(box6 + 3)};
    Ray r{v3(ray6), v3(ray6 + 3)};
    float tt = 0;
    bool ok = slab(b, r, tt, semantics);
    *t = tt;
    return ok ? 1 : 0;
}
int32_t xo_tri_test(const XnartTriangle* s, const float ray6[6], float* u, float* v, float* d) {
    Tri t;
    t.v1 = v3(s->v1); t.v2 = v3(s->v2); t.v3 = v3(s->v3); t.sn = v3(s->surfaceNormal);
    Ray r{v3(ray6), v3(ray6 + 3)};
    return triTest(r, t, *u, *v, *d);
}
uint32_t xo_pack_color(const float rgb[3], int32_t semantics) { return packColor(v3(rgb), semantics); }
// channel pack of an already multiplied value t = v*255 (tests half-even on exact .5 products)
uint32_t xo_round_channel_premultiplied(float t) {
    if (std::isnan(t)) return 0u;
    if (t < 0.0f) return 0u;
    if (t > 255.0f) return 255u;
    return (uint32_t)std::nearbyint((double)t);
}
void xo_unpack_color(uint32_t c, float rgb[3]) { V3 v = unpackColor(c); rgb[0] = v.x; rgb[1] = v.y; rgb[2] = v.z; }
void xo_matrix_invert(const float in[16], float out[16]) { M4 r = invert(toM4(in)); std::memcpy(out, r.m, 64); }
void xo_matrix_multiply(const float a[16], const float b[16], float out[16]) { M4 r = multiply(toM4(a), toM4(b)); std::memcpy(out, r.m, 64); }
void xo_create_look_at(const float eye[3], const float target[3], const float up[3], float out[16]) {
    M4 r = createLookAt(v3(eye), v3(target), v3(up)); std::memcpy(out, r.m, 64);
}
void xo_create_perspective_fov(float fov, float aspect, float zn, float zf, float out[16]) {
    M4 r = createPerspectiveFov(fov, aspect, zn, zf); std::memcpy(out, r.m, 64);
}
void xo_lookup_uv(const XnartMaterial* m, const float uv[2], int32_t addressMode, int32_t filtering, float rgb[3]) {
    Material mat;
    mat.useTexture = true; mat.texW = m->texW; mat.texH = m->texH;
    mat.argb.assign(m->argb, m->argb + (size_t)m->texW * (size_t)m->texH);
    V3 c = lookupUV(mat, V2{uv[0], uv[1]}, addressMode, filtering);
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}
void xo_light_for_fragment(const XnartLight* l, const float pos[3], const float normal[3], float rgb[3]) {
    Light L;
    L.kind = l->kind; L.position = v3(l->position); L.direction = v3(l->direction); L.color = v3(l->color);
    L.intensity = l->intensity; L.decayExponent = l->decayExponent; L.spotAngle = l->spotAngle;
    L.angleCosine = (float)std::cos((double)(L.spotAngle * 0.5f));
    V3 c = lightForFragment(L, v3(pos), v3(normal));
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}
void xo_normalize(const float in[3], float out[3]) { V3 r = normalize(v3(in)); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void xo_reflect(const float v[3], const float n[3], float out[3]) { V3 r = reflect(v3(v), v3(n)); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void xo_transform(const float p[3], const float m[16], float out[3]) { V3 r = transform(v3(p), toM4(m)); out[0] = r.x; out[1] = r.y; out[2] = r.z; }

}  // extern "C"
