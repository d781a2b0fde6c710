/* TEST INFRASTRUCTURE ONLY — SURVEY.md §7 "step 0": can the GPU box create a headless OpenGL context?
 *
 * No GL/EGL headers exist in the image, so the handful of entry points and enums are declared by hand and the
 * libraries are dlopen'ed. Prints what it finds and exits 0 if an OpenGL >= 3.3 context was made current (then the
 * reference's GLSL files could be executed unmodified to pin oracle/efo_map.cpp), 1 otherwise.
 * Build: gcc -O1 -o oracle/_ref/egl_probe oracle/egl_probe.c -ldl
 */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

typedef void* EGLDisplay;
typedef void* EGLConfig;
typedef void* EGLContext;
typedef void* EGLSurface;
typedef void* EGLDeviceEXT;
typedef int EGLint;
typedef unsigned int EGLBoolean;
typedef unsigned int EGLenum;

#define EGL_PLATFORM_DEVICE_EXT 0x313F
#define EGL_OPENGL_API 0x30A2
#define EGL_OPENGL_BIT 0x0008
#define EGL_RENDERABLE_TYPE 0x3040
#define EGL_SURFACE_TYPE 0x3033
#define EGL_PBUFFER_BIT 0x0001
#define EGL_NONE 0x3038
#define EGL_CONTEXT_MAJOR_VERSION 0x3098
#define EGL_CONTEXT_MINOR_VERSION 0x30FB
#define EGL_WIDTH 0x3057
#define EGL_HEIGHT 0x3056
#define EGL_VENDOR 0x3053
#define EGL_VERSION 0x3054
#define EGL_EXTENSIONS 0x3055
#define GL_VENDOR 0x1F00
#define GL_RENDERER 0x1F01
#define GL_VERSION 0x1F02

typedef void* (*PFN_getproc)(const char*);

int main(void) {
  const char* names[] = {"libEGL.so.1", "libEGL.so", "libEGL_nvidia.so.0", NULL};
  void* h = NULL;
  for (int i = 0; names[i] && !h; ++i) {
    h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    printf("dlopen(%s) -> %s\n", names[i], h ? "ok" : dlerror());
  }
  if (!h) {
    printf("RESULT: no EGL library\n");
    return 1;
  }
  PFN_getproc getproc = (PFN_getproc)dlsym(h, "eglGetProcAddress");
  if (!getproc) {
    printf("RESULT: no eglGetProcAddress\n");
    return 1;
  }
#define EGLFN(ret, name, args) ret(*name) args = (ret(*) args)dlsym(h, #name); if (!name) name = (ret(*) args)getproc(#name);
  EGLFN(EGLDisplay, eglGetDisplay, (void*));
  EGLFN(EGLBoolean, eglInitialize, (EGLDisplay, EGLint*, EGLint*));
  EGLFN(const char*, eglQueryString, (EGLDisplay, EGLint));
  EGLFN(EGLBoolean, eglBindAPI, (EGLenum));
  EGLFN(EGLBoolean, eglChooseConfig, (EGLDisplay, const EGLint*, EGLConfig*, EGLint, EGLint*));
  EGLFN(EGLContext, eglCreateContext, (EGLDisplay, EGLConfig, EGLContext, const EGLint*));
  EGLFN(EGLSurface, eglCreatePbufferSurface, (EGLDisplay, EGLConfig, const EGLint*));
  EGLFN(EGLBoolean, eglMakeCurrent, (EGLDisplay, EGLSurface, EGLSurface, EGLContext));
  EGLFN(EGLint, eglGetError, (void));
  EGLBoolean (*eglQueryDevicesEXT)(EGLint, EGLDeviceEXT*, EGLint*) = (EGLBoolean(*)(EGLint, EGLDeviceEXT*, EGLint*))getproc("eglQueryDevicesEXT");
  EGLDisplay (*eglGetPlatformDisplayEXT)(EGLenum, void*, const EGLint*) = (EGLDisplay(*)(EGLenum, void*, const EGLint*))getproc("eglGetPlatformDisplayEXT");
  const char* client_ext = eglQueryString ? eglQueryString(NULL, EGL_EXTENSIONS) : NULL;
  printf("client extensions: %s\n", client_ext ? client_ext : "(none)");
  EGLDisplay dpy = NULL;
  if (eglQueryDevicesEXT && eglGetPlatformDisplayEXT) {
    EGLDeviceEXT devs[16];
    EGLint nd = 0;
    if (eglQueryDevicesEXT(16, devs, &nd)) {
      printf("eglQueryDevicesEXT: %d devices\n", nd);
      for (int i = 0; i < nd && !dpy; ++i) {
        EGLDisplay d = eglGetPlatformDisplayEXT(EGL_PLATFORM_DEVICE_EXT, devs[i], NULL);
        EGLint ma = 0, mi = 0;
        if (d && eglInitialize(d, &ma, &mi)) {
          printf("device %d: EGL %d.%d vendor=%s\n", i, ma, mi, eglQueryString(d, EGL_VENDOR));
          dpy = d;
        } else {
          printf("device %d: eglInitialize failed (0x%x)\n", i, eglGetError());
        }
      }
    } else {
      printf("eglQueryDevicesEXT failed (0x%x)\n", eglGetError());
    }
  } else {
    printf("no EGL_EXT_device_enumeration / platform_device\n");
  }
  if (!dpy) {
    EGLDisplay d = eglGetDisplay(NULL);
    EGLint ma = 0, mi = 0;
    if (d && eglInitialize(d, &ma, &mi)) {
      printf("default display: EGL %d.%d vendor=%s\n", ma, mi, eglQueryString(d, EGL_VENDOR));
      dpy = d;
    } else {
      printf("default display: eglInitialize failed (0x%x)\n", eglGetError());
    }
  }
  if (!dpy) {
    printf("RESULT: no EGL display\n");
    return 1;
  }
  if (!eglBindAPI(EGL_OPENGL_API)) {
    printf("RESULT: eglBindAPI(OPENGL) failed (0x%x)\n", eglGetError());
    return 1;
  }
  const EGLint cfg_attr[] = {EGL_SURFACE_TYPE, EGL_PBUFFER_BIT, EGL_RENDERABLE_TYPE, EGL_OPENGL_BIT, EGL_NONE};
  EGLConfig cfg = NULL;
  EGLint ncfg = 0;
  if (!eglChooseConfig(dpy, cfg_attr, &cfg, 1, &ncfg) || ncfg < 1) {
    printf("RESULT: no pbuffer/OpenGL config (0x%x)\n", eglGetError());
    return 1;
  }
  const EGLint ctx_attr[] = {EGL_CONTEXT_MAJOR_VERSION, 3, EGL_CONTEXT_MINOR_VERSION, 3, EGL_NONE};
  EGLContext ctx = eglCreateContext(dpy, cfg, NULL, ctx_attr);
  if (!ctx) {
    printf("RESULT: eglCreateContext(3.3) failed (0x%x)\n", eglGetError());
    return 1;
  }
  const EGLint pb_attr[] = {EGL_WIDTH, 16, EGL_HEIGHT, 16, EGL_NONE};
  EGLSurface surf = eglCreatePbufferSurface(dpy, cfg, pb_attr);
  if (!eglMakeCurrent(dpy, surf, surf, ctx)) {
    printf("RESULT: eglMakeCurrent failed (0x%x)\n", eglGetError());
    return 1;
  }
  const unsigned char* (*glGetString)(unsigned int) = (const unsigned char* (*)(unsigned int))getproc("glGetString");
  if (!glGetString) {
    printf("RESULT: context current but no glGetString\n");
    return 1;
  }
  printf("GL_VENDOR=%s\nGL_RENDERER=%s\nGL_VERSION=%s\n", glGetString(GL_VENDOR), glGetString(GL_RENDERER), glGetString(GL_VERSION));
  printf("glTransformFeedbackVaryings=%p glDrawTransformFeedback=%p glTexStorage2D=%p\n", getproc("glTransformFeedbackVaryings"),
         getproc("glDrawTransformFeedback"), getproc("glTexStorage2D"));
  printf("RESULT: headless OpenGL context OK\n");
  return 0;
}
