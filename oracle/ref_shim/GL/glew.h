// Stand-in (NOT OpenGL; test infrastructure): the two enumerators SiftGPU is handed (feature extraction, out of scope).
#pragma once
#ifndef GL_BGR
#define GL_BGR 0x80E0
#define GL_UNSIGNED_BYTE 0x1401
#endif
