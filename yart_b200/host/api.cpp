// api.cpp — host layer of libyart_b200.so: the C mirror of yart's Scene loading, Camera and
// Renderer API (ys_* / yr_* in include/yart_cuda.h) on top of the device layer (yc_*).
//
//   ys_scene_load     ↔ gltf::load + Mesh/BVH construction      src/gltf/gltf.cpp:319-358, src/core/mesh.hpp:54-61
//   ys_camera_make    ↔ Camera ctor + moveAndLookAt             src/core/camera.hpp:25-59, 77-136
//   yr_render / yr_render_sync / yr_abort / yr_wait
//                     ↔ Renderer::render / renderSync / abort / wait   src/core/renderer.hpp:56-95
//   wave schedule     ↔ TileRenderer::renderImpl / finishTile   src/cpu/tile-renderer.hpp:118-124, 264-288
// Compiled with g++ -ffp-contract=off (derived constants must round like the reference's).
#include <atomic>
#include <chrono>
#include <mutex>
#include <new>
#include <thread>
#include <cstring>
#include <vector>

#include "scene.hpp"

using namespace yartb;

namespace yartb {
bool loadGlb(const std::string& path, ysc::SceneDesc& out, std::string* err);
bool writePpm(const std::string& path, const float* rgba, uint32_t w, uint32_t h);
bool decodeTextureForTest(const uint8_t* png, size_t len, uint32_t type, int C, const int* channels, ysc::TextureDesc& out,
                          std::string& err);
}  // namespace yartb

struct ys_scene {
  HostScene host;
};

static thread_local std::string g_ysError;

extern "C" const char* ys_last_error(void) { return g_ysError.c_str(); }

extern "C" int ys_scene_load_bvh(const char* path, uint32_t bvhKind, ys_scene** out);
extern "C" int ys_scene_load(const char* path, ys_scene** out) { return ys_scene_load_bvh(path, YS_BVH_SAH, out); }

extern "C" int ys_scene_load_bvh(const char* path, uint32_t bvhKind, ys_scene** out) {
  if (!path || !out || bvhKind > YS_BVH_MEDIAN_SPLIT) return YC_ERR_INVALID;
  *out = nullptr;
  ysc::SceneDesc d;
  std::string err;
  if (!ysc::load(path, d, &err)) {
    g_ysError = err;
    return YC_ERR_IO;
  }
  ys_scene* s = new (std::nothrow) ys_scene();
  if (!s) return YC_ERR_INVALID;
  s->host.bvhKind = bvhKind;
  if (!s->host.build(d, &err)) {
    g_ysError = err;
    delete s;
    return YC_ERR_INVALID;
  }
  *out = s;
  return YC_OK;
}

// main.cpp:81-84: `ImageInfiniteLight(radius, &hdri); scene->addLight(...)` after gltf::load
static void appendEnv(ysc::SceneDesc& d, const YsEnvLight* env) {
  if (!env || !env->rgb || env->width < 2 || env->height < 2) return;
  ysc::TextureDesc t;
  t.channels = 3, t.isFloat = 1, t.type = ysc::LinearRGB, t.width = env->width, t.height = env->height;
  t.f32.assign(env->rgb, env->rgb + size_t(env->width) * env->height * 3);
  d.textures.push_back(std::move(t));
  ysc::LightDesc l;
  l.type = ysc::ImageInfiniteT;
  l.hdrTex = int32_t(d.textures.size()) - 1;
  l.sceneRadius = env->sceneRadius;
  l.hasTransform = env->hasTransform;
  if (env->hasTransform) memcpy(l.m, env->transform, sizeof l.m);
  d.lights.push_back(l);
}

extern "C" int ys_glb_convert(const char* glbPath, const char* yscPath, const YsEnvLight* env) {
  if (!glbPath || !yscPath) return YC_ERR_INVALID;
  ysc::SceneDesc d;
  std::string err;
  if (!loadGlb(glbPath, d, &err)) {
    g_ysError = err;
    return YC_ERR_IO;
  }
  appendEnv(d, env);
  if (!ysc::save(yscPath, d)) {
    g_ysError = std::string("cannot write ") + yscPath;
    return YC_ERR_IO;
  }
  return YC_OK;
}

extern "C" int ys_scene_load_glb(const char* path, const YsEnvLight* env, ys_scene** out) {
  if (!path || !out) return YC_ERR_INVALID;
  *out = nullptr;
  ysc::SceneDesc d;
  std::string err;
  if (!loadGlb(path, d, &err)) {
    g_ysError = err;
    return YC_ERR_IO;
  }
  appendEnv(d, env);
  ys_scene* s = new (std::nothrow) ys_scene();
  if (!s) return YC_ERR_INVALID;
  if (!s->host.build(d, &err)) {
    g_ysError = err;
    delete s;
    return YC_ERR_INVALID;
  }
  *out = s;
  return YC_OK;
}

extern "C" int ys_decode_texture(const void* png, size_t len, uint32_t type, uint32_t nChannels, const int32_t* channels,
                                 uint8_t* out, size_t outBytes, uint32_t* width, uint32_t* height) {
  if (!png || !channels || nChannels < 1 || nChannels > 4) return YC_ERR_INVALID;
  ysc::TextureDesc t;
  std::string err;
  int ch[4] = {0, 1, 2, 3};
  for (uint32_t i = 0; i < nChannels; i++) ch[i] = channels[i];
  if (!decodeTextureForTest(static_cast<const uint8_t*>(png), len, type, int(nChannels), ch, t, err)) {
    g_ysError = err;
    return YC_ERR_IO;
  }
  if (width) *width = t.width;
  if (height) *height = t.height;
  if (out) {
    if (outBytes < t.u8.size()) return YC_ERR_INVALID;
    memcpy(out, t.u8.data(), t.u8.size());
  }
  return YC_OK;
}

extern "C" int ys_write_ppm(const char* path, const float* rgba, uint32_t width, uint32_t height) {
  if (!path || !rgba || !width || !height) return YC_ERR_INVALID;
  return writePpm(path, rgba, width, height) ? YC_OK : YC_ERR_IO;
}

extern "C" void ys_scene_destroy(ys_scene* s) { delete s; }
extern "C" const YcScene* ys_scene_flat(const ys_scene* s) { return s ? &s->host.flat : nullptr; }
extern "C" double ys_scene_build_ms(const ys_scene* s) { return s ? s->host.buildMs : 0.0; }

extern "C" int ys_scene_bvh(const ys_scene* s, uint32_t mesh, const void** nodes, uint32_t* nNodes,
                            const uint32_t** indices, uint32_t* nTris) {
  if (!s || mesh >= s->host.refBvh.size()) return YC_ERR_INVALID;
  const BvhBuildResult& b = s->host.refBvh[mesh];
  if (nodes) *nodes = b.nodes.data();
  if (nNodes) *nNodes = uint32_t(b.nodes.size());
  if (indices) *indices = b.indices.data();
  if (nTris) *nTris = uint32_t(b.indices.size());
  return YC_OK;
}

// Camera(imageSize, focalLength, fNumber) with the default 36x24 sensor, then moveAndLookAt
// (camera.hpp:131-136 → setDirection :98-105 → calcDerivedProperties :25-59).
extern "C" int ys_camera_make(uint32_t width, uint32_t height, float focalLength, float fNumber, const float position[3],
                              const float target[3], const float up[3], float exposure, uint32_t apertureSides,
                              YcCamera* out) {
  if (!out || !position || !target || width == 0 || height == 0) return YC_ERR_INVALID;
  const float sensorX = 36.0f, sensorY = 24.0f;
  const f3 pos(position), forward = f3(target) - f3(position);
  f3 mUp(0.0f, 1.0f, 0.0f);
  if (up && length2(f3(up)) != 0.0f) mUp = f3(up);
  const float aspect = float(width) / float(height);
  const float sensorAspect = sensorX / sensorY;
  const float croppedSensorHeight = sensorX / rmax(sensorAspect, aspect);
  const float focusDistance = length(forward);
  const float vh = focusDistance * croppedSensorHeight / focalLength;
  const float vw = vh * aspect;
  mUp = normalized(mUp);
  const f3 w = normalized(-forward);
  const f3 u = cross(mUp, w);
  const f3 v = cross(w, u);
  const Frame frame(w, u);
  const f3 viewportU = u * vw;
  const f3 viewportV = (-v) * vh;
  const f3 viewportTopLeft = pos - w * focusDistance - (viewportU + viewportV) * 0.5f;
  const f3 pixelDeltaU = viewportU / float(width);
  const f3 pixelDeltaV = viewportV / float(height);
  const f3 topLeftPixel = viewportTopLeft + (pixelDeltaU + pixelDeltaV) * 0.5f;
  const float apertureRadius = fNumber ? (focalLength / 2000.0f) / fNumber : 0.0f;
  auto put = [](float* d, f3 s) { d[0] = s.x, d[1] = s.y, d[2] = s.z; };
  put(out->position, pos);
  put(out->topLeftPixel, topLeftPixel);
  put(out->pixelDeltaU, pixelDeltaU);
  put(out->pixelDeltaV, pixelDeltaV);
  put(out->frameX, frame.x);
  put(out->frameY, frame.y);
  put(out->frameZ, frame.z);
  out->apertureRadius = apertureRadius;
  out->apertureSides = apertureSides;
  out->exposure = exposure;
  return YC_OK;
}

// ---------------------------------------------------------------------------------------
// renderer
// ---------------------------------------------------------------------------------------
struct yr_renderer {
  YrSettings s{};
  const ys_scene* scene = nullptr;
  YcCamera cam{};
  yc_ctx* ctx = nullptr;
  bool sceneUploaded = false;
  std::string err;
  yr_wave_callback cb = nullptr;
  void* cbUser = nullptr;
  std::thread worker;
  std::atomic<bool> stop{false};
  std::mutex m;
  YrRenderData last{};
  int lastRc = YC_OK;
};

static int rfail(yr_renderer* r, int rc, const std::string& msg) {
  r->err = msg;
  return rc;
}

extern "C" int yr_create(const YrSettings* settings, const ys_scene* scene, const YcCamera* camera, yr_renderer** out) {
  if (!settings || !camera || !out) return YC_ERR_INVALID;
  *out = nullptr;
  if (settings->width == 0 || settings->height == 0 || settings->samples == 0 || settings->tileSize == 0)
    return YC_ERR_INVALID;
  yr_renderer* r = new (std::nothrow) yr_renderer();
  if (!r) return YC_ERR_INVALID;
  r->s = *settings;
  r->scene = scene;
  r->cam = *camera;
  YcOptions o{};
  o.maxDepth = settings->maxDepth;
  o.integrator = settings->integrator;
  o.scrambler = settings->scrambler;
  o.sampler = settings->sampler;
  o.traversal = settings->traversal;
  int rc = yc_create(settings->device, &o, &r->ctx);
  if (rc != YC_OK) {
    delete r;
    return rc;
  }
  *out = r;
  return YC_OK;
}

extern "C" void yr_destroy(yr_renderer* r) {
  if (!r) return;
  r->stop = true;
  if (r->worker.joinable()) r->worker.join();
  yc_destroy(r->ctx);
  delete r;
}

extern "C" int yr_set_wave_callback(yr_renderer* r, yr_wave_callback cb, void* user) {
  if (!r) return YC_ERR_INVALID;
  r->cb = cb;
  r->cbUser = user;
  return YC_OK;
}

extern "C" yc_ctx* yr_context(yr_renderer* r) { return r ? r->ctx : nullptr; }
extern "C" const char* yr_last_error(const yr_renderer* r) { return r ? r->err.c_str() : "null renderer"; }

static int renderBlocking(yr_renderer* r, YrRenderData* out) {
  using clock = std::chrono::high_resolution_clock;
  if (!r->scene) return rfail(r, YC_ERR_NO_SCENE, "no scene (reference: `if (!scene) return;`, integrator.cpp:6)");
  int rc;
  if (!r->sceneUploaded) {
    if ((rc = yc_upload_scene(r->ctx, ys_scene_flat(r->scene))) != YC_OK) return rfail(r, rc, yc_last_error(r->ctx));
    r->sceneUploaded = true;
  }
  if ((rc = yc_set_camera(r->ctx, &r->cam)) != YC_OK) return rfail(r, rc, yc_last_error(r->ctx));
  YcFrameDesc f{};
  f.width = r->s.width, f.height = r->s.height;
  f.totalSamples = r->s.samples, f.tileSize = r->s.tileSize;
  for (int k = 0; k < 3; k++) f.background[k] = r->s.background[k];
  f.tonemap = r->s.tonemap, f.estimator = r->s.estimator;
  f.shardIndex = r->s.shardIndex, f.shardCount = r->s.shardCount ? r->s.shardCount : 1;
  if ((rc = yc_begin_frame(r->ctx, &f)) != YC_OK) return rfail(r, rc, yc_last_error(r->ctx));

  // wave schedule: tile-renderer.hpp:120-124 (reset) and :264-288 (advance)
  const uint64_t total = r->s.samples;
  uint64_t remaining = total, wave = 0;
  uint64_t waveSamples = std::min<uint64_t>(r->s.firstWaveSamples ? r->s.firstWaveSamples : total, total);
  const uint64_t maxWave = r->s.maxWaveSamples ? r->s.maxWaveSamples : total;
  const auto t0 = clock::now();
  YrRenderData data{};
  data.totalSamples = total;
  uint64_t raysBefore = 0;
  while (waveSamples > 0 && !r->stop) {
    const auto w0 = clock::now();
    const YcRect full{0, 0, r->s.width, r->s.height};
    rc = yc_render_wave(r->ctx, full, uint32_t(total - remaining), uint32_t(waveSamples), uint32_t(total - remaining));
    if (rc != YC_OK) return rfail(r, rc, yc_last_error(r->ctx));
    remaining -= waveSamples;
    YcStats st{};
    yc_resolve(r->ctx, nullptr, nullptr, &st);
    const auto now = clock::now();
    data.samplesTaken = total - remaining;
    data.totalRays = st.raysReference;
    data.totalTimeMs = std::chrono::duration<double, std::milli>(now - t0).count();
    if (r->cb) {
      YrWaveData wd{wave, waveSamples, st.raysReference - raysBefore, std::chrono::duration<double, std::milli>(now - w0).count()};
      r->cb(&data, &wd, r->cbUser);
    }
    raysBefore = st.raysReference;
    const uint64_t next = (wave > 0 || waveSamples > 1) ? std::min<uint64_t>(waveSamples * 2, maxWave) : 1;
    waveSamples = std::min<uint64_t>(next, remaining);
    wave++;
  }
  {
    std::lock_guard<std::mutex> lk(r->m);
    r->last = data;
  }
  if (out) *out = data;
  return YC_OK;
}

extern "C" int yr_render_sync(yr_renderer* r, YrRenderData* out) {
  if (!r) return YC_ERR_INVALID;
  if (r->worker.joinable()) r->worker.join();
  r->stop = false;
  return r->lastRc = renderBlocking(r, out);
}

extern "C" int yr_render(yr_renderer* r) {
  if (!r) return YC_ERR_INVALID;
  r->stop = true;  // Renderer::render aborts a render in progress first (tile-renderer.hpp:40-43)
  if (r->worker.joinable()) r->worker.join();
  r->stop = false;
  r->worker = std::thread([r] { r->lastRc = renderBlocking(r, nullptr); });
  return YC_OK;
}

extern "C" int yr_abort(yr_renderer* r) {
  if (!r) return YC_ERR_INVALID;
  r->stop = true;
  if (r->worker.joinable()) r->worker.join();
  return YC_OK;
}

extern "C" int yr_wait(yr_renderer* r) {
  if (!r) return YC_ERR_INVALID;
  if (r->worker.joinable()) r->worker.join();
  return r->lastRc;
}

extern "C" int yr_write_ppm(yr_renderer* r, const char* path) {
  if (!r || !path) return YC_ERR_INVALID;
  std::vector<float> ldr(size_t(r->s.width) * r->s.height * 4);
  int rc = yc_resolve(r->ctx, nullptr, ldr.data(), nullptr);
  if (rc != YC_OK) return rfail(r, rc, yc_last_error(r->ctx));
  return writePpm(path, ldr.data(), r->s.width, r->s.height) ? YC_OK : rfail(r, YC_ERR_IO, std::string("cannot write ") + path);
}

extern "C" int yr_read(yr_renderer* r, float* hdrRGBA, float* ldrRGBA, YcStats* stats) {
  if (!r) return YC_ERR_INVALID;
  int rc = yc_resolve(r->ctx, hdrRGBA, ldrRGBA, stats);
  if (rc != YC_OK) return rfail(r, rc, yc_last_error(r->ctx));
  return YC_OK;
}
