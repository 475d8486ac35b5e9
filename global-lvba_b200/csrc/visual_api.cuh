// visual_api.cuh — boundary B2 (placeholder until the kernels land)
#pragma once
#include "runtime.cuh"
struct lvba_visual_problem { int dummy; };
extern "C" {
void lvba_visual_default_opts(lvba_visual_opts* o) {
  if (!o) return;
  o->max_iter = 50; o->initial_radius = 1e4; o->max_radius = 1e16; o->min_radius = 1e-32;
  o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32; o->min_relative_decrease = 1e-3;
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->jacobi_scaling = 1; o->device = -1; o->verbose = 0;
}
}
