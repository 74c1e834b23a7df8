/*
 * gl_stub.cpp -- TEST INFRASTRUCTURE (never part of the product).  A headless stand-in for the OpenGL / GLFW / GLEW /
 * CUDA-GL-interop entry points the reference's display shell calls (src/main.cpp, src/preview.cpp, src/glslUtility.cpp), so
 * that those files can be compiled UNMODIFIED, where they lie, and linked against libptb200.so's pathtraceInit / pathtraceFree /
 * pathtrace (oracle/Makefile target `glshell` -> oracle/_ref/ref_gl_shell).  There is no GL stack in this
 * image: windows, textures and shaders are no-ops, and the pixel buffer object the app maps every frame
 * (src/main.cpp:128-137 cudaGLMapBufferObject) is backed by plain device memory, so the library's sendImageToPBO path runs for
 * real.  The app then saves its PNG and exits by itself after ITERATIONS frames (src/main.cpp:141-146).
 */
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdio>

extern "C" {

/* ---- GLFW ---- */
static int g_window;
int glfwInit(void) { return 1; }
void glfwTerminate(void) {}
void *glfwCreateWindow(int, int, const char *, void *, void *) { return &g_window; }
void glfwDestroyWindow(void *) {}
void glfwMakeContextCurrent(void *) {}
void glfwPollEvents(void) {}
void glfwSwapBuffers(void *) {}
void glfwSetWindowTitle(void *, const char *) {}
void glfwSetWindowShouldClose(void *, int) {}
int glfwWindowShouldClose(void *) { return 0; }
void *glfwSetErrorCallback(void *) { return 0; }
void *glfwSetKeyCallback(void *, void *) { return 0; }
void *glfwSetCursorPosCallback(void *, void *) { return 0; }
void *glfwSetMouseButtonCallback(void *, void *) { return 0; }

/* ---- GL 1.1 entry points ---- */
void glBindTexture(unsigned, unsigned) {}
void glClear(unsigned) {}
void glDeleteTextures(int, const unsigned *) {}
void glDrawElements(unsigned, int, unsigned, const void *) {}
void glGenTextures(int n, unsigned *t) { for (int i = 0; i < n; ++i) t[i] = 1; }
void glTexImage2D(unsigned, int, int, int, int, int, unsigned, unsigned, const void *) {}
void glTexParameteri(unsigned, unsigned, int) {}
void glTexSubImage2D(unsigned, int, int, int, int, int, unsigned, unsigned, const void *) {}

/* ---- GLEW: the extension entry points are function-pointer variables ---- */
unsigned char glewExperimental = 0;
unsigned glewInit(void) { return 0; /* GLEW_OK */ }
static size_t g_pboBytes = 0;
static void *g_pboDev = NULL;
static void stubVoid() {}
static unsigned stubOne() { return 1; }
static int stubLoc() { return 0; }
static void stubGen(int n, unsigned *ids) { for (int i = 0; i < n; ++i) ids[i] = 1; }
static void stubBufferData(unsigned, ptrdiff_t size, const void *, unsigned) { g_pboBytes = (size_t)size; }   /* the PBO is the last buffer sized (preview.cpp initPBO) */
static void stubGetiv(unsigned, unsigned pname, int *v) { *v = (pname == 0x8B84 /* GL_INFO_LOG_LENGTH */) ? 0 : 1; }
#define GLEW_FN(name, fn) void *name = (void *)(fn)
GLEW_FN(__glewActiveTexture, stubVoid); GLEW_FN(__glewAttachShader, stubVoid); GLEW_FN(__glewBindAttribLocation, stubVoid);
GLEW_FN(__glewBindBuffer, stubVoid); GLEW_FN(__glewBufferData, stubBufferData); GLEW_FN(__glewCompileShader, stubVoid);
GLEW_FN(__glewCreateProgram, stubOne); GLEW_FN(__glewCreateShader, stubOne); GLEW_FN(__glewDeleteBuffers, stubVoid);
GLEW_FN(__glewEnableVertexAttribArray, stubVoid); GLEW_FN(__glewGenBuffers, stubGen); GLEW_FN(__glewGetProgramInfoLog, stubVoid);
GLEW_FN(__glewGetProgramiv, stubGetiv); GLEW_FN(__glewGetShaderInfoLog, stubVoid); GLEW_FN(__glewGetShaderiv, stubGetiv);
GLEW_FN(__glewGetUniformLocation, stubLoc); GLEW_FN(__glewLinkProgram, stubVoid); GLEW_FN(__glewShaderSource, stubVoid);
GLEW_FN(__glewUniform1i, stubVoid); GLEW_FN(__glewUseProgram, stubVoid); GLEW_FN(__glewVertexAttribPointer, stubVoid);

/* ---- CUDA <-> GL interop (cuda_gl_interop.h, deprecated API the reference uses): the PBO is plain device memory ---- */
cudaError_t cudaGLSetGLDevice(int device) { return cudaSetDevice(device); }
cudaError_t cudaGLRegisterBufferObject(unsigned) { return cudaSuccess; }
cudaError_t cudaGLUnregisterBufferObject(unsigned) { if (g_pboDev) cudaFree(g_pboDev); g_pboDev = NULL; return cudaSuccess; }
cudaError_t cudaGLMapBufferObject(void **devPtr, unsigned) {
    if (!g_pboDev && g_pboBytes) { if (cudaMalloc(&g_pboDev, g_pboBytes) != cudaSuccess) { g_pboDev = NULL; cudaGetLastError(); } }
    *devPtr = g_pboDev;
    return cudaSuccess;
}
cudaError_t cudaGLUnmapBufferObject(unsigned) { return cudaSuccess; }

} /* extern "C" */
