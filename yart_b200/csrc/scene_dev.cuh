// scene_dev.cuh — device-resident scene (pointers into HBM) shared by all kernels.
//
// HBM layout (all read-only during a render):
//   nodes      YcNode[nNodes]            scene graph, DFS pre-order (144 B each)
//   meshes     YcMesh[nMeshes]
//   bvhNodes   float4[4 * nBvhNodes]     inner BVH2 nodes, both child boxes inlined (64 B, 4 × LDG.128)
//   bvhTris    float4[3 * nBvhTris]      leaf-ordered triangles, positions pre-gathered (48 B, 3 × LDG.128)
//   wideNodes  float4[4 * nWide]         the same tree collapsed to 4-wide nodes with quantised boxes (64 B, wide_bvh.cuh), scenes without alpha
//   positions/normals/tangents/uvs       per-vertex attributes (shade-time gathers)
//   primIndices/primMaterial/primLight   per-primitive, original order
//   materials, textures, texels, lights, envDist, light-sampler tables, LUTs
#pragma once
#include "../../include/yart_cuda.h"
#include "dmath.cuh"

namespace yb {

// Per-mesh entry into the wide (4-ary) node array built by yc_upload_scene (wide_bvh.cuh).
struct WideMesh {
  uint32_t rootRef;     // like YcMesh::rootRef, relative to this mesh's first WideNode
  uint32_t nodeOffset;  // first WideNode of this mesh
};

struct DScene {
  const YcNode* nodes;
  uint32_t nNodes;
  const int32_t* nodePath;  // [nNodes][YC_MAX_NODE_DEPTH]: ancestors of node i from the root (level 0) down to i
  const YcMesh* meshes;
  uint32_t nMeshes;
  const float4* bvhNodes;
  const float4* bvhTris;
  const float4* wideNodes;      // float4[4 * nWideNodes]: the collapsed 4-wide nodes with quantised child boxes (64 B each), null when not built
  const WideMesh* wideMeshes;   // [nMeshes]
  const float* positions;
  const float* normals;
  const float* tangents;
  const float* uvs;
  const uint32_t* primIndices;
  const uint32_t* primMaterial;
  const int32_t* primLight;
  const YcMaterial* materials;
  const YcTexture* textures;
  const uint8_t* texU8;
  const float* texF32;
  const YcLight* lights;
  uint32_t nLights;
  const float* envDist;
  const uint32_t* infLights;
  uint32_t nInf;
  const uint32_t* areaLights;
  uint32_t nArea;
  const float* powerCdf;
  float totalPower;
  uint32_t uniformLights;  // YB_RNG_SAMPLERS build: UniformLightSampler instead of PowerLightSampler (light-sampler.cpp:11-31)
  const float* lut;
  int hasAlpha;
};

// Closest-hit record the wavefront carries between extend and shade (what is needed to
// rebuild the reference's Hit, src/cpu/hit.hpp:8-17, at shade time).
struct HitRec {
  float t;
  float u, v;     // barycentrics of p1, p2 (Hit::tg before testMesh overwrites it)
  uint32_t prim;  // Hit::idx
  int32_t node;   // scene-graph node whose mesh was hit, -1 = miss
  uint32_t backSide;
};

}  // namespace yb
