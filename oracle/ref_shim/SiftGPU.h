// Stand-in (NOT SiftGPU; test infrastructure): feature extraction and matching are outside the hot path;
// extractAndMatchFeaturesGPU compiles against these declarations and reports "not supported" if it is ever called.
#pragma once
#include <vector>
class SiftGPU {
 public:
  struct SiftKeypoint { float x, y, s, o; };
  enum { SIFTGPU_NOT_SUPPORTED = 0, SIFTGPU_PARTIAL_SUPPORTED = 1, SIFTGPU_FULL_SUPPORTED = 2 };
  void ParseParam(int, const char**) {}
  void ParseParam(int, char**) {}
  int CreateContextGL() { return SIFTGPU_NOT_SUPPORTED; }
  int RunSIFT(int, int, const void*, unsigned, unsigned) { return 0; }
  int GetFeatureNum() { return 0; }
  void GetFeatureVector(SiftKeypoint*, float*) {}
};
class SiftMatchGPU {
 public:
  explicit SiftMatchGPU(int = 4096) {}
  int VerifyContextGL() { return 0; }
  void SetMaxSift(int) {}
  void SetDescriptors(int, int, const float*, int = -1) {}
  void SetDescriptors(int, int, const unsigned char*, int = -1) {}
  int GetSiftMatch(int, int (*)[2], float = 0.7f, float = 0.8f, int = 1) { return 0; }
  int GetSiftMatch(int, unsigned (*)[2], float = 0.7f, float = 0.8f, int = 1) { return 0; }
};
