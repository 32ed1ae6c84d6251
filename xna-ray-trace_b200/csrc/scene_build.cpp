// scene_build.cpp — host-side octree construction for libxnart.so.
//
// Produces, for arbitrary input, exactly the tree the reference builds:
//   MeshOctree.Build/BuildTree/SplitCuboid            RayTracerTypeLibrary/MeshOctree.cs:56-96, 204-236
//   OctreeSpatialManager.Build/BuildTree/SplitCuboid  RayTraceProject/RayTraceProject/Spatial/OctreeSpatialManager.cs:64-113, 218-248
// but organised for large inputs: the tree is grown level by level (so nodes come out in breadth
// order with the 8 children of a node consecutive — the device layout), and each split makes ONE
// pass over the parent's items computing an 8-bit child-membership mask per item (parallel over
// items) followed by an order-preserving scatter, instead of the reference's eight filter passes.
// Item order inside every child list equals the parent's order, as in the reference
// (MeshOctree.cs:224-231), and leaf lists are finally laid out in depth-first leaf order.
// All arithmetic is binary32, unfused (compile with -ffp-contract=off).
#include "scene_build.h"

#include <algorithm>
#include <cfloat>
#include <cstring>

#include "device_types.h"

namespace xn {

namespace {

struct BNode {
    HBox b;
    int firstChild = -1;
    int depth = 0;
    std::vector<int> items;
};

// Child boxes of SplitCuboid: size = (Max - Min) / 2f  (Vector3 / float multiplies by 1f/2f),
// pos = Min + (size.X*i, size.Y*j, size.Z*k), box = (pos, pos + size).
struct Halves {
    float lo[3][2];  // [axis][half] -> child Min on that axis
    float hi[3][2];  // [axis][half] -> child Max on that axis
};

static Halves halvesOf(const HBox& p) {
    Halves h;
    const float half = 1.0f / 2.0f;
    for (int a = 0; a < 3; a++) {
        float size = (p.mx[a] - p.mn[a]) * half;
        for (int s = 0; s < 2; s++) {
            float pos = p.mn[a] + (size * (float)s);
            h.lo[a][s] = pos;
            h.hi[a][s] = pos + size;
        }
    }
    return h;
}

static HBox childBox(const Halves& h, int c) {
    const int s[3] = {(c >> 2) & 1, (c >> 1) & 1, c & 1};
    HBox b;
    for (int a = 0; a < 3; a++) { b.mn[a] = h.lo[a][s[a]]; b.mx[a] = h.hi[a][s[a]]; }
    return b;
}

// BoundingBox.Contains(Vector3) != Disjoint, per axis and half: Min <= p <= Max.
static inline unsigned pointMask(const Halves& h, const float* p) {
    unsigned ax[3];
    for (int a = 0; a < 3; a++) {
        unsigned m = 0;
        if (!(h.lo[a][0] > p[a] || p[a] > h.hi[a][0])) m |= 1u;
        if (!(h.lo[a][1] > p[a] || p[a] > h.hi[a][1])) m |= 2u;
        ax[a] = m;
    }
    unsigned mask = 0;
    for (int c = 0; c < 8; c++)
        if (((ax[0] >> ((c >> 2) & 1)) & 1u) && ((ax[1] >> ((c >> 1) & 1)) & 1u) && ((ax[2] >> (c & 1)) & 1u)) mask |= 1u << c;
    return mask;
}

// BoundingBox.Intersects(BoundingBox), inclusive; `ob` may have unsorted corners (Q8).
static inline unsigned boxMask(const Halves& h, const HBox& ob) {
    unsigned ax[3];
    for (int a = 0; a < 3; a++) {
        unsigned m = 0;
        for (int s = 0; s < 2; s++)
            if (!(ob.mx[a] < h.lo[a][s] || ob.mn[a] > h.hi[a][s])) m |= 1u << s;
        ax[a] = m;
    }
    unsigned mask = 0;
    for (int c = 0; c < 8; c++)
        if (((ax[0] >> ((c >> 2) & 1)) & 1u) && ((ax[1] >> ((c >> 1) & 1)) & 1u) && ((ax[2] >> (c & 1)) & 1u)) mask |= 1u << c;
    return mask;
}

template <class MaskFn>
static int growTree(const HBox& rootBox, size_t nItems, int threshold, const MaskFn& maskOf, FlatTree& out,
                    std::string& err) {
    std::vector<BNode> nodes;
    nodes.reserve(1024);
    nodes.emplace_back();
    nodes[0].b = rootBox;
    nodes[0].items.resize(nItems);
    for (size_t i = 0; i < nItems; i++) nodes[0].items[i] = (int)i;
    int maxDepth = 0;
    std::vector<unsigned char> masks;
    for (size_t n = 0; n < nodes.size(); n++) {
        if ((long long)nodes[n].items.size() <= (long long)threshold) continue;  // BuildTree: Count > itemTreshold
        if (nodes[n].depth >= kBuilderDepthGuard) {
            err = "octree would recurse without bound (more than itemThreshold items share one point); "
                  "the reference overflows its stack here (MeshOctree.cs:84-96)";
            return XNART_ERR_OCTREE_DEPTH;
        }
        const Halves h = halvesOf(nodes[n].b);
        {
            // All eight children identical to the parent (a box collapsed to a point): every level
            // would reproduce this node eight-fold for ever.  The reference recurses without bound
            // here; report it now instead of growing 8^depth nodes up to the depth guard.
            bool same = true;
            for (int a = 0; a < 3; a++)
                for (int s = 0; s < 2; s++)
                    same = same && h.lo[a][s] == nodes[n].b.mn[a] && h.hi[a][s] == nodes[n].b.mx[a];
            if (same) {
                err = "octree would recurse without bound (more than itemThreshold items share one point); "
                      "the reference overflows its stack here (MeshOctree.cs:84-96)";
                return XNART_ERR_OCTREE_DEPTH;
            }
            if (nodes.size() > ((size_t)1 << 28)) {
                err = "octree exceeds 2^28 nodes";
                return XNART_ERR_UNSUPPORTED;
            }
        }
        std::vector<int> items;
        items.swap(nodes[n].items);  // inner nodes' containingObjects are never read by traversal
        const long long cnt = (long long)items.size();
        masks.resize((size_t)cnt);
#pragma omp parallel for schedule(static) if (cnt > 65536)
        for (long long i = 0; i < cnt; i++) masks[(size_t)i] = (unsigned char)maskOf(h, items[(size_t)i]);
        size_t counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (long long i = 0; i < cnt; i++) {
            unsigned m = masks[(size_t)i];
            for (int c = 0; c < 8; c++) counts[c] += (m >> c) & 1u;
        }
        const int first = (int)nodes.size();
        const int depth = nodes[n].depth + 1;
        nodes[n].firstChild = first;
        maxDepth = std::max(maxDepth, depth);
        for (int c = 0; c < 8; c++) {
            nodes.emplace_back();
            BNode& ch = nodes.back();
            ch.b = childBox(h, c);
            ch.depth = depth;
            ch.items.reserve(counts[c]);
        }
        for (long long i = 0; i < cnt; i++) {
            unsigned m = masks[(size_t)i];
            while (m) {
                int c = __builtin_ctz(m);
                m &= m - 1;
                nodes[(size_t)first + (size_t)c].items.push_back(items[(size_t)i]);
            }
        }
    }

    const size_t N = nodes.size();
    out = FlatTree();
    out.depth = maxDepth;
    out.lo.resize(3 * N); out.hi.resize(3 * N); out.w0.resize(N); out.w1.resize(N);
    // leaf lists in depth-first (child 0..7) leaf order
    std::vector<int> stack;
    stack.push_back(0);
    while (!stack.empty()) {
        int n = stack.back();
        stack.pop_back();
        if (nodes[(size_t)n].firstChild >= 0) {
            for (int c = 7; c >= 0; c--) stack.push_back(nodes[(size_t)n].firstChild + c);
        } else {
            out.w0[(size_t)n] = (int)out.refs.size();
            out.w1[(size_t)n] = (int)nodes[(size_t)n].items.size();
            out.refs.insert(out.refs.end(), nodes[(size_t)n].items.begin(), nodes[(size_t)n].items.end());
        }
    }
    // boxes + conservative Max (children always have larger indices than their parent)
    std::vector<float> cons(3 * N);
    for (size_t n = 0; n < N; n++)
        for (int a = 0; a < 3; a++) {
            out.lo[3 * n + a] = nodes[n].b.mn[a];
            out.hi[3 * n + a] = nodes[n].b.mx[a];
            cons[3 * n + a] = nodes[n].b.mx[a];
        }
    int consIdx = 0;
    for (size_t n = 0; n < N; n++)
        if (nodes[n].firstChild >= 0) {
            out.w0[n] = nodes[n].firstChild;
            out.w1[n] = -1 - consIdx++;
        }
    out.consMax.resize(3 * (size_t)consIdx);
    for (size_t k = N; k-- > 0;) {
        if (nodes[k].firstChild < 0) continue;
        for (int c = 0; c < 8; c++) {
            size_t ch = (size_t)nodes[k].firstChild + (size_t)c;
            for (int a = 0; a < 3; a++) {
                if (cons[3 * ch + a] > cons[3 * k + a]) cons[3 * k + a] = cons[3 * ch + a];
                // a child's Min is Min + size*i with size >= 0, so it can never undershoot: traversal relies on it
                if (nodes[ch].b.mn[a] < nodes[k].b.mn[a]) {
                    err = "internal: child Min below parent Min";
                    return XNART_ERR_UNSUPPORTED;
                }
            }
        }
        size_t ci = (size_t)(-1 - out.w1[k]);
        for (int a = 0; a < 3; a++) out.consMax[3 * ci + a] = cons[3 * k + a];
    }
    return XNART_OK;
}

}  // namespace

int buildMeshOctree(const XnartTriangle* tris, int64_t nTri, int itemThreshold, FlatTree& out, std::string& err) {
    // MeshOctree.Build: box starts at (float.MaxValue, float.MinValue) and is shrunk with
    // Vector3.Min/Max ((a < b) ? a : b / (a > b) ? a : b) over v1, v2, v3 of every triangle.
    HBox box;
    for (int a = 0; a < 3; a++) { box.mn[a] = FLT_MAX; box.mx[a] = -FLT_MAX; }
    for (int64_t i = 0; i < nTri; i++) {
        const float* vs[3] = {tris[i].v1, tris[i].v2, tris[i].v3};
        for (int k = 0; k < 3; k++)
            for (int a = 0; a < 3; a++) {
                box.mn[a] = (box.mn[a] < vs[k][a]) ? box.mn[a] : vs[k][a];
                box.mx[a] = (box.mx[a] > vs[k][a]) ? box.mx[a] : vs[k][a];
            }
    }
    // vertex containment (MeshOctree.cs:226-228): the triangle goes to every child that contains v1, v2 or v3
    auto maskOf = [tris](const Halves& h, int i) {
        return pointMask(h, tris[i].v1) | pointMask(h, tris[i].v2) | pointMask(h, tris[i].v3);
    };
    return growTree(box, (size_t)nTri, itemThreshold, maskOf, out, err);
}

static inline void transformPoint(const float* p, const float* m, float* o) {
    for (int c = 0; c < 3; c++) o[c] = (((p[0] * m[0 + c]) + (p[1] * m[4 + c])) + (p[2] * m[8 + c])) + m[12 + c];
}

int buildSceneOctree(std::vector<HostObj>& objs, int itemThreshold, FlatTree& out, std::string& err) {
    HBox box;
    for (int a = 0; a < 3; a++) { box.mn[a] = 0.0f; box.mx[a] = 0.0f; }  // new BoundingBox(Vector3.Zero, Vector3.Zero)
    for (size_t i = 0; i < objs.size(); i++) {
        HostObj& o = objs[i];
        float tmn[3], tmx[3];
        transformPoint(o.localBox.mn, o.world, tmn);
        transformPoint(o.localBox.mx, o.world, tmx);
        HBox ob;  // re-sorted corners for the root bounds (OctreeSpatialManager.cs:83-85)
        for (int a = 0; a < 3; a++) {
            ob.mn[a] = (tmn[a] < tmx[a]) ? tmn[a] : tmx[a];
            ob.mx[a] = (tmn[a] > tmx[a]) ? tmn[a] : tmx[a];
        }
        if (i == 0) box = ob;
        else
            for (int a = 0; a < 3; a++) {  // BoundingBox.CreateMerged
                box.mn[a] = (box.mn[a] < ob.mn[a]) ? box.mn[a] : ob.mn[a];
                box.mx[a] = (box.mx[a] > ob.mx[a]) ? box.mx[a] : ob.mx[a];
            }
        // SceneObject.WorldBoundingBox (SceneObject.cs:195-196): both corners transformed, NOT re-sorted
        std::memcpy(o.worldBox.mn, tmn, sizeof tmn);
        std::memcpy(o.worldBox.mx, tmx, sizeof tmx);
    }
    const std::vector<HostObj>* po = &objs;
    auto maskOf = [po](const Halves& h, int i) { return boxMask(h, (*po)[(size_t)i].worldBox); };
    return growTree(box, objs.size(), itemThreshold, maskOf, out, err);
}

void packTriangles(const XnartTriangle* tris, int64_t nTri, std::vector<float>& triI, std::vector<float>& triS) {
    triI.resize((size_t)nTri * 16);
    triS.resize((size_t)nTri * 20);
#pragma omp parallel for schedule(static) if (nTri > 65536)
    for (int64_t i = 0; i < nTri; i++) {
        const XnartTriangle& t = tris[i];
        float* I = &triI[(size_t)i * 16];
        float* S = &triS[(size_t)i * 20];
        for (int a = 0; a < 3; a++) {
            I[a] = t.surfaceNormal[a];
            I[4 + a] = t.v1[a];
            I[8 + a] = t.v2[a] - t.v1[a];   // Edge1 (RayExtensions.cs:54)
            I[12 + a] = t.v3[a] - t.v1[a];  // Edge2 (RayExtensions.cs:55)
        }
        I[3] = I[7] = I[11] = I[15] = 0.0f;
        S[0] = t.uv1[0]; S[1] = t.uv1[1]; S[2] = t.uv2[0]; S[3] = t.uv2[1];
        S[4] = t.uv3[0]; S[5] = t.uv3[1]; S[6] = t.n1[0]; S[7] = t.n1[1];
        S[8] = t.n1[2]; S[9] = t.n2[0]; S[10] = t.n2[1]; S[11] = t.n2[2];
        S[12] = t.n3[0]; S[13] = t.n3[1]; S[14] = t.n3[2]; S[15] = t.color[0];
        S[16] = t.color[1]; S[17] = t.color[2]; S[18] = t.color[3]; S[19] = 0.0f;
    }
}

}  // namespace xn
