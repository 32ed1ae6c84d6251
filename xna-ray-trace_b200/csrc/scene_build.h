// scene_build.h — host-side construction of the flattened octrees and triangle records.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/xnart.h"

namespace xn {

struct HBox { float mn[3], mx[3]; };

// One octree in its final (device) numbering: breadth order, children of a node consecutive.
struct FlatTree {
    std::vector<float> lo;        // 3 floats per node (Min)
    std::vector<float> hi;        // 3 floats per node (Max)
    std::vector<int> w0, w1;      // leaf: (refStart, count)   internal: (firstChild, -1 - consIndex)
    std::vector<float> consMax;   // 3 floats per internal node
    std::vector<int> refs;        // leaf item lists in depth-first leaf order
    int depth = 0;
    size_t nodeCount() const { return w0.size(); }
};

struct HostMesh {
    std::vector<float> triI;      // 16 floats per triangle (n,0 | v1,0 | E1,0 | E2,0)
    std::vector<float> triS;      // 20 floats per triangle (see device_types.h)
    int64_t nTri = 0;
    HBox box{};                   // Mesh.MeshBoundingBox
    float reflectiveness = 0, refractionIndex = 0;
    int flags = 0;
    int texW = 0, texH = 0;
    std::vector<uint32_t> texels;
    FlatTree tree;
};

struct HostObj {
    std::vector<int> meshIds;
    float world[16], inv[16];
    HBox localBox{};
    HBox worldBox{};              // corners transformed without re-sorting (SceneObject.cs:195-196, Q8)
};

// MeshOctree.Build + BuildTree + SplitCuboid (MeshOctree.cs:56-96, 204-236).  Returns XNART_OK or
// XNART_ERR_OCTREE_DEPTH when a node deeper than kBuilderDepthGuard would still have to split.
int buildMeshOctree(const XnartTriangle* tris, int64_t nTri, int itemThreshold, FlatTree& out, std::string& err);

// OctreeSpatialManager.Build + SplitCuboid (OctreeSpatialManager.cs:64-113, 218-248).
int buildSceneOctree(std::vector<HostObj>& objs, int itemThreshold, FlatTree& out, std::string& err);

// Repack Triangle records (Triangle.cs:14-25) into the intersection / shading arrays.
void packTriangles(const XnartTriangle* tris, int64_t nTri, std::vector<float>& triI, std::vector<float>& triS);

}  // namespace xn
