// adapter_test.cpp — TEST INFRASTRUCTURE.  Compiles integration/wavefront-renderer.hpp (the yart::Renderer adapter a
// maintainer would add to teofum/yart) against the UNMODIFIED reference sources and renders the same yart::Scene —
// built through the reference's public API by ref_driver.cpp's buildScene — with the reference's own
// cpu::TileRenderer and with yart::cuda::WavefrontRenderer, in one process, and compares the frames bit for bit.
// Linked twice by oracle/Makefile: against tests/hostsim/libyart_hostsim.so (the product sources compiled for the CPU;
// runs anywhere) and against yart_b200/libyart_b200.so (CUDA; runs on the GPU box).
//
//   adapter_* <scene.ysc> <out.bin> [w= h= spp= first= max= tile= maxdepth= tonemap=agx|golden|punchy|none pos= target=
//             focal= fnum= exposure= sides= traversal=0|1|2 devices=0,1,...]
// Prints one JSON line; out.bin = header + the adapter's LDR and HDR frames.
#define main ref_driver_main
#include "ref_driver.cpp"
#undef main

#include "../integration/wavefront-renderer.hpp"

static bool sameBits(const Buffer& a, const Buffer& b, size_t* differing) {
  size_t n = 0;
  for (uint32_t y = 0; y < a.height(); y++)
    for (uint32_t x = 0; x < a.width(); x++) {
      const float4 p = a(x, y), q = b(x, y);
      for (size_t k = 0; k < 4; k++) {
        uint32_t u, v;
        memcpy(&u, &p[k], 4), memcpy(&v, &q[k], 4);
        if (u != v && !(std::isnan(p[k]) && std::isnan(q[k]))) n++;
      }
    }
  *differing = n;
  return n == 0;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s scene.ysc out.bin [key=value ...]\n", argv[0]);
    return 1;
  }
  Args a(argc, argv, 3);
  ysc::SceneDesc d;
  std::string err;
  if (!ysc::load(argv[1], d, &err)) {
    fprintf(stderr, "%s\n", err.c_str());
    return 2;
  }
  RefScene rs = buildScene(d);
  const uint32_t w = uint32_t(a.num("w", 64)), h = uint32_t(a.num("h", 64));
  Camera cam = makeCamera(a, w, h);
  g_maxDepth = uint32_t(a.num("maxdepth", 30));
  tonemap::AgX agx;
  const std::string tm = a.str("tonemap", "agx");
  if (tm == "golden") agx.look = tonemap::AgX::golden;
  if (tm == "punchy") agx.look = tonemap::AgX::punchy;
  const tonemap::Tonemap* tonemapper = tm == "none" ? nullptr : &agx;
  const float3 bg = a.vec3("bg", {0, 0, 0});

  // ---- the reference's renderer --------------------------------------------------------------------
  cpu::TileRenderer<RefSampler, DepthIntegrator> ref(Buffer(w, h), cam);
  ref.scene = rs.scene.get();
  ref.samples = uint32_t(a.num("spp", 16));
  ref.firstWaveSamples = uint32_t(a.num("first", ref.samples));
  ref.maxWaveSamples = uint32_t(a.num("max", ref.samples));
  ref.tileSize = uint32_t(a.num("tile", 64));
  if (a.has("threads")) ref.threadCount = uint32_t(a.num("threads", 1));
  ref.backgroundColor = bg;
  ref.tonemapper = tonemapper;
  size_t refWaves = 0, refTiles = 0, refDone = 0;
  ref.onRenderWaveComplete = [&](Renderer::RenderData, Renderer::WaveData) { refWaves++; };
  // (the reference adds tile ray counts into m_totalRays outside its mutex, tile-renderer.hpp:217-218: the sum of the
  // per-tile counts, handed over under m_bufferMutex, is the exact figure — see ref_driver.cpp)
  uint64_t refTileRays = 0;
  ref.onRenderTileComplete = [&](Renderer::RenderData, Renderer::TileData t) { refTiles++, refTileRays += t.rays; };
  ref.onRenderComplete = [&](Renderer::RenderData) { refDone++; };
  auto ra = ref.renderSync();
  ra.totalRays = refTileRays;

  // ---- the adapter, on the same yart::Scene and Camera objects ----------------------------------------
  cuda::WavefrontRenderer gpu(Buffer(w, h), cam);
  gpu.scene = rs.scene.get();
  gpu.samples = ref.samples, gpu.firstWaveSamples = ref.firstWaveSamples, gpu.maxWaveSamples = ref.maxWaveSamples;
  gpu.tileSize = ref.tileSize;
  gpu.backgroundColor = bg;
  gpu.tonemapper = tonemapper;
  gpu.maxDepth = g_maxDepth;
  gpu.traversal = uint32_t(a.num("traversal", YC_TRAVERSAL_REFERENCE_ORDER));
  if (a.has("devices")) {
    gpu.devices.clear();
    std::stringstream ss(a.str("devices", "0"));
    for (std::string tok; std::getline(ss, tok, ',');) gpu.devices.push_back(atoi(tok.c_str()));
  }
  size_t gpuWaves = 0, gpuTiles = 0, gpuDone = 0, gpuAborted = 0;
  uint64_t waveRaySum = 0;
  gpu.onRenderWaveComplete = [&](Renderer::RenderData, Renderer::WaveData wd) { gpuWaves++, waveRaySum += wd.rays; };
  gpu.onRenderTileComplete = [&](Renderer::RenderData, Renderer::TileData) { gpuTiles++; };
  gpu.onRenderComplete = [&](Renderer::RenderData) { gpuDone++; };
  gpu.onRenderAborted = [&](Renderer::RenderData) { gpuAborted++; };
  const auto rb = gpu.renderSync();
  if (rb.samplesTaken != rb.totalSamples) {
    fprintf(stderr, "adapter render failed: %s\n", gpu.lastError());
    return 3;
  }
  size_t diffLdr = 0, diffHdr = 0, diffAsync = 0;
  sameBits(ra.buffer, rb.buffer, &diffLdr);
  Buffer hdr(w, h);
  gpu.readHdr(hdr);
  sameBits(ref.m_hdrBuffer, hdr, &diffHdr);  // TileRenderer's private accumulation buffer (-fno-access-control)
  const uint64_t gpuRays = rb.totalRays;
  const size_t wavesSync = gpuWaves, tilesSync = gpuTiles, doneSync = gpuDone;
  Buffer first(w, h);
  for (uint32_t y = 0; y < h; y++)
    for (uint32_t x = 0; x < w; x++) first(x, y) = rb.buffer(x, y);

  // asynchronous interface: render() + wait() gives the same frame again; abort() ends a render early (or not at
  // all if it already finished) and exactly one of the two completion callbacks fires per render
  gpu.render();
  gpu.wait();
  sameBits(first, rb.buffer, &diffAsync);
  const size_t doneAfterAsync = gpuDone;
  gpu.render();
  gpu.abort();
  gpu.wait();
  const bool oneCompletion = gpuDone + gpuAborted == doneAfterAsync + 1;

  Writer out(argv[2]);
  out.put(w);
  out.put(h);
  out.put(uint64_t(gpuRays));
  for (uint32_t y = 0; y < h; y++)
    for (uint32_t x = 0; x < w; x++) out.putn(first(x, y).data(), 4);
  for (uint32_t y = 0; y < h; y++)
    for (uint32_t x = 0; x < w; x++) out.putn(hdr(x, y).data(), 4);
  printf("{\"ldr_words_differing\": %zu, \"hdr_words_differing\": %zu, \"async_words_differing\": %zu, \"rays_reference\": %llu, "
         "\"rays_adapter\": %llu, \"wave_ray_sum\": %llu, \"waves_reference\": %zu, \"waves_adapter\": %zu, \"tiles_reference\": %zu, "
         "\"tiles_adapter\": %zu, \"done_reference\": %zu, \"done_adapter\": %zu, \"aborted_adapter\": %zu, \"one_completion_per_render\": %s, "
         "\"devices\": %zu}\n",
         diffLdr, diffHdr, diffAsync, (unsigned long long) ra.totalRays, (unsigned long long) gpuRays, (unsigned long long) waveRaySum,
         refWaves, wavesSync, refTiles, tilesSync, refDone, doneSync, gpuAborted, oneCompletion ? "true" : "false", gpu.devices.size());
  return 0;
}
