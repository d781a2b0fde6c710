/* TEST INFRASTRUCTURE ONLY — can the Mesa/llvmpipe libGL bundled with Nsight Compute give us an OpenGL 3.3 context without an X
 * server (through oracle/gl/fake_x11.c)? Prints GL_VERSION / GL_RENDERER, exit 0 on success. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#define __USE_GNU
#include <signal.h>
#include <ucontext.h>
#include <execinfo.h>
#include <string.h>
#include <unistd.h>
static void on_segv(int sig, siginfo_t* si, void* uc_) {
  ucontext_t* uc = (ucontext_t*)uc_;
  fprintf(stderr, "SIGSEGV addr=%p rip=%p rdi=%p rsi=%p rax=%p\n", si->si_addr, (void*)uc->uc_mcontext.gregs[REG_RIP], (void*)uc->uc_mcontext.gregs[REG_RDI], (void*)uc->uc_mcontext.gregs[REG_RSI], (void*)uc->uc_mcontext.gregs[REG_RAX]);
  void* bt[32]; int n = backtrace(bt, 32); backtrace_symbols_fd(bt, n, 2);
  FILE* f = fopen("/proc/self/maps", "r"); char line[512]; while (fgets(line, sizeof line, f)) if (strstr(line, "libGL") && strstr(line, "r-xp")) fputs(line, stderr);
  _exit(139);
}

typedef void* (*getproc_t)(const char*);
int main(int argc, char** argv) {
  setvbuf(stdout, NULL, _IONBF, 0);
  struct sigaction sa; memset(&sa, 0, sizeof sa); sa.sa_sigaction = on_segv; sa.sa_flags = SA_SIGINFO; sigaction(SIGSEGV, &sa, NULL);
  const char* lib = argc > 1 ? argv[1] : "libGL.so.1";
  void* x11 = dlopen("libX11.so.6", RTLD_NOW | RTLD_GLOBAL);
  if (!x11) { printf("no libX11 stand-in: %s\n", dlerror()); return 1; }
  void* (*fake_display)(void) = (void* (*)(void))dlsym(x11, "fake_x11_display");
  unsigned long (*fake_window)(int, int) = (unsigned long (*)(int, int))dlsym(x11, "fake_x11_window");
  if (!fake_display) { printf("libX11.so.6 is a real Xlib, not the stand-in\n"); return 1; }
  printf("stand-in loaded, opening %s\n", lib);
  void* gl = dlopen(lib, RTLD_NOW | RTLD_GLOBAL);
  printf("libGL loaded: %p\n", gl);
  if (!gl) { printf("dlopen(%s): %s\n", lib, dlerror()); return 1; }
  getproc_t getproc = (getproc_t)dlsym(gl, "glXGetProcAddressARB");
  void* dpy = fake_display();
  void** (*chooseFB)(void*, int, const int*, int*) = (void** (*)(void*, int, const int*, int*))dlsym(gl, "glXChooseFBConfig");
  const int fb_attr[] = {0x8010 /*GLX_DRAWABLE_TYPE*/, 0x1 /*WINDOW*/, 0x8011 /*RENDER_TYPE*/, 0x1 /*RGBA*/, 8 /*RED*/, 8, 9, 8, 10, 8, 12 /*DEPTH*/, 24, 0};
  int n = 0;
  printf("calling glXChooseFBConfig (%p) dpy=%p\n", (void*)chooseFB, dpy);
  void** cfgs = chooseFB(dpy, 0, fb_attr, &n);
  printf("glXChooseFBConfig -> %d configs\n", n);
  if (!cfgs || n < 1) return 1;
  void* (*createAttribs)(void*, void*, void*, int, const int*) = (void* (*)(void*, void*, void*, int, const int*))getproc("glXCreateContextAttribsARB");
  printf("glXCreateContextAttribsARB = %p\n", (void*)createAttribs);
  void* ctx = NULL;
  if (createAttribs) {
    const int ca[] = {0x2091 /*MAJOR*/, 3, 0x2092 /*MINOR*/, 3, 0x9126 /*PROFILE_MASK*/, 0x1 /*CORE*/, 0};
    ctx = createAttribs(dpy, cfgs[0], NULL, 1, ca);
  }
  if (!ctx) {
    void* (*createNew)(void*, void*, int, void*, int) = (void* (*)(void*, void*, int, void*, int))dlsym(gl, "glXCreateNewContext");
    ctx = createNew(dpy, cfgs[0], 0x8014 /*GLX_RGBA_TYPE*/, NULL, 1);
    printf("fell back to glXCreateNewContext -> %p\n", ctx);
  }
  if (!ctx) { printf("RESULT: no context\n"); return 1; }
  unsigned long win = fake_window(64, 64);
  int (*makeCurrent)(void*, unsigned long, void*) = (int (*)(void*, unsigned long, void*))dlsym(gl, "glXMakeCurrent");
  if (!makeCurrent(dpy, win, ctx)) { printf("RESULT: glXMakeCurrent failed\n"); return 1; }
  const unsigned char* (*getString)(unsigned) = (const unsigned char* (*)(unsigned))getproc("glGetString");
  printf("GL_VENDOR=%s\nGL_RENDERER=%s\nGL_VERSION=%s\nGLSL=%s\n", getString(0x1F00), getString(0x1F01), getString(0x1F02), getString(0x8B8C));
  printf("RESULT: OpenGL context OK\n");
  return 0;
}
