// scene.hpp — host mirror of yart's Scene and its flattening into the GPU layout (YcScene).
//
// HostScene::build performs, on a neutral SceneDesc, the steps the reference performs through
// its C++ API (src/gltf/gltf.cpp:319-358 order: materials → meshes (+ SAH BVH) → node tree →
// lights), computing every derived constant with the reference's arithmetic:
//   Node bounds            src/core/scene.hpp:17-22, 54-58 (+ transform.hpp:75-90 padding)
//   AreaLight area / power src/core/light.cpp:16-40
//   ImageInfiniteLight     src/core/light.cpp:137-196, math/sampling.hpp:118-152
//   PowerLightSampler      src/core/light-sampler.cpp:32-50
//   ParametricBSDF ctor    src/bsdf/parametric.cpp:49-66
// and then lays everything out as flat arrays (include/yart_cuda.h).
#pragma once
#include <string>
#include <vector>

#include "../../include/yart_cuda.h"
#include "bvh_build.hpp"
#include "hmath.hpp"
#include "scene_desc.hpp"

namespace yartb {

struct HostScene {
  // flattened storage (owned)
  std::vector<YcNode> nodes;
  std::vector<YcMesh> meshes;
  std::vector<YcBvhNode> bvhNodes;
  std::vector<YcBvhTri> bvhTris;
  std::vector<float> positions, normals, tangents, uvs;
  std::vector<uint32_t> primIndices, primMaterial;
  std::vector<int32_t> primLight;
  std::vector<YcMaterial> materials;
  std::vector<YcTexture> textures;
  std::vector<uint8_t> texelsU8;
  std::vector<float> texelsF32;
  std::vector<YcLight> lights;
  std::vector<float> envDist;
  std::vector<uint32_t> infiniteLights, areaLights;
  std::vector<float> lightPowerCdf;
  std::vector<float> lutTables;
  float totalPower = 0;
  // reference-layout BVHs kept for parity tests
  std::vector<BvhBuildResult> refBvh;
  double buildMs = 0;
  uint32_t deviceBuilds = 0;  // meshes whose SAH BVH was built on the GPU (yc_build_bvh_sah)
  YcScene flat{};

  uint32_t bvhKind = 0;  // YS_BVH_* (Mesh::BVHType, mesh.hpp:17)
  bool build(const ysc::SceneDesc& d, std::string* err);
  bool loadLuts(std::string* err);
};

int setBuildDevice(int device);  // ys_set_build_device

}  // namespace yartb
