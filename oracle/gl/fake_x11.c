/* TEST INFRASTRUCTURE ONLY — a stand-in for libX11.so.6 / libXext.so.6 so that the Xlib-GLX build of Mesa (llvmpipe) that ships
 * with Nsight Compute in this image (/opt/nvidia/nsight-compute/.../Mesa/libGL.so.1, Mesa 18.1.9) can create an OFF-SCREEN OpenGL
 * 3.3 context without an X server. It exports exactly the Xlib entry points that library imports and answers them for one fake
 * display: one screen, one 24-bit TrueColor visual, no MIT-SHM, windows / pixmaps as plain ids with a size. Nothing is ever
 * drawn to a "window": the GL harness renders into framebuffer objects and reads them back, as the reference does.
 * The structure layouts are those of <X11/Xlib.h> / <X11/Xutil.h> (public, unchanged since X11R6), restated here because the
 * image has no X11 headers. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned long XID;
typedef XID Window, Drawable, Pixmap, Colormap, VisualID, Font;
typedef char* XPointer;
typedef int Bool;
typedef struct _XGC* GC;

typedef struct _XExtData XExtData;
typedef struct {
  XExtData* ext_data;
  VisualID visualid;
  int c_class;
  unsigned long red_mask, green_mask, blue_mask;
  int bits_per_rgb;
  int map_entries;
} Visual;
typedef struct {
  int depth;
  int nvisuals;
  Visual* visuals;
} Depth;
struct _XDisplay;
typedef struct {
  XExtData* ext_data;
  struct _XDisplay* display;
  Window root;
  int width, height;
  int mwidth, mheight;
  int ndepths;
  Depth* depths;
  int root_depth;
  Visual* root_visual;
  GC default_gc;
  Colormap cmap;
  unsigned long white_pixel;
  unsigned long black_pixel;
  int max_maps, min_maps;
  int backing_store;
  Bool save_unders;
  long root_input_mask;
} Screen;
typedef struct {
  XExtData* ext_data;
  int depth;
  int bits_per_pixel;
  int scanline_pad;
} ScreenFormat;
typedef struct _XDisplay {
  XExtData* ext_data;
  void* private1;
  int fd;
  int private2;
  int proto_major_version;
  int proto_minor_version;
  char* vendor;
  XID private3;
  XID private4;
  XID private5;
  int private6;
  XID (*resource_alloc)(struct _XDisplay*);
  int byte_order;
  int bitmap_unit;
  int bitmap_pad;
  int bitmap_bit_order;
  int nformats;
  ScreenFormat* pixmap_format;
  int private8;
  int release;
  void *private9, *private10;
  int qlen;
  unsigned long last_request_read;
  unsigned long request;
  XPointer private11;
  XPointer private12;
  XPointer private13;
  XPointer private14;
  unsigned max_request_size;
  void* db;
  int (*private15)(struct _XDisplay*);
  char* display_name;
  int default_screen;
  int nscreens;
  Screen* screens;
  unsigned long motion_buffer;
  unsigned long private16;
  int min_keycode;
  int max_keycode;
  XPointer private17;
  XPointer private18;
  int private19;
  char* xdefaults;
  char pad[4096]; /* the private tail of the real structure */
} Display;

typedef struct {
  Visual* visual;
  VisualID visualid;
  int screen;
  int depth;
  int c_class;
  unsigned long red_mask, green_mask, blue_mask;
  int colormap_size;
  int bits_per_rgb;
} XVisualInfo;

typedef struct _XImage {
  int width, height;
  int xoffset;
  int format;
  char* data;
  int byte_order;
  int bitmap_unit;
  int bitmap_bit_order;
  int bitmap_pad;
  int depth;
  int bytes_per_line;
  int bits_per_pixel;
  unsigned long red_mask, green_mask, blue_mask;
  XPointer obdata;
  struct funcs {
    struct _XImage* (*create_image)(Display*, Visual*, unsigned, int, int, char*, unsigned, unsigned, int, int);
    int (*destroy_image)(struct _XImage*);
    unsigned long (*get_pixel)(struct _XImage*, int, int);
    int (*put_pixel)(struct _XImage*, int, int, unsigned long);
    struct _XImage* (*sub_image)(struct _XImage*, int, int, unsigned, unsigned);
    int (*add_pixel)(struct _XImage*, long);
  } f;
} XImage;

typedef struct {
  int x, y;
  int width, height;
  int border_width;
  int depth;
  Visual* visual;
  Window root;
  int c_class;
  int bit_gravity;
  int win_gravity;
  int backing_store;
  unsigned long backing_planes;
  unsigned long backing_pixel;
  Bool save_under;
  Colormap colormap;
  Bool map_installed;
  int map_state;
  long all_event_masks;
  long your_event_mask;
  long do_not_propagate_mask;
  Bool override_redirect;
  Screen* screen;
} XWindowAttributes;

typedef struct {
  int extension;
  int major_opcode;
  int first_event;
  int first_error;
} XExtCodes;

#define TrueColor 4
#define ZPixmap 2
#define LSBFirst 0

static Visual g_visual = {NULL, 0x21, TrueColor, 0xff0000, 0x00ff00, 0x0000ff, 8, 256};
static Depth g_depth = {24, 1, &g_visual};
static Screen g_screen;
static ScreenFormat g_format = {NULL, 24, 32, 32};
static Display g_display;
static int g_init = 0;
static XID g_next_id = 0x400000;
/* drawables: id -> size */
static struct { XID id; int w, h; } g_draw[256];
static int g_ndraw = 0;

/* the harness calls this instead of XOpenDisplay */
Display* fake_x11_display(void) {
  if (!g_init) {
    memset(&g_display, 0, sizeof(g_display));
    memset(&g_screen, 0, sizeof(g_screen));
    g_screen.display = &g_display;
    g_screen.root = 0x100;
    g_screen.width = 1280;
    g_screen.height = 1024;
    g_screen.mwidth = 340;
    g_screen.mheight = 270;
    g_screen.ndepths = 1;
    g_screen.depths = &g_depth;
    g_screen.root_depth = 24;
    g_screen.root_visual = &g_visual;
    g_screen.cmap = 0x20;
    g_screen.white_pixel = 0xffffff;
    g_screen.max_maps = g_screen.min_maps = 1;
    g_display.fd = -1;
    g_display.proto_major_version = 11;
    g_display.vendor = (char*)"fake_x11 (efusion-b200 oracle)";
    g_display.byte_order = LSBFirst;
    g_display.bitmap_unit = 32;
    g_display.bitmap_pad = 32;
    g_display.bitmap_bit_order = LSBFirst;
    g_display.nformats = 1;
    g_display.pixmap_format = &g_format;
    g_display.release = 12101000;
    g_display.max_request_size = 65535;
    g_display.display_name = (char*)":fake";
    g_display.default_screen = 0;
    g_display.nscreens = 1;
    g_display.screens = &g_screen;
    g_init = 1;
  }
  return &g_display;
}
/* a fake window of a given size to make current against (never drawn to) */
Window fake_x11_window(int w, int h) {
  XID id = g_next_id++;
  if (g_ndraw < 256) {
    g_draw[g_ndraw].id = id;
    g_draw[g_ndraw].w = w;
    g_draw[g_ndraw].h = h;
    g_ndraw++;
  }
  return id;
}
static void draw_size(XID d, int* w, int* h) {
  *w = 64;
  *h = 64;
  for (int i = 0; i < g_ndraw; ++i)
    if (g_draw[i].id == d) {
      *w = g_draw[i].w;
      *h = g_draw[i].h;
    }
}

void (*_XLockMutex_fn)(void*) = NULL;
void (*_XUnlockMutex_fn)(void*) = NULL;
void* _Xglobal_lock = NULL;

/* Mesa's GLX registers itself by walking / prepending to Display::ext_procs, a PRIVATE member of struct _XDisplay (Xlibint.h:
 * at byte 0x140 on LP64, after xdefaults, scratch_buffer, scratch_length, ext_number); struct _XExten = { next, XExtCodes codes,
 * nine hook pointers (close_display at 0x48), char* name at 0x60, ... }. */
typedef struct _XExten {
  struct _XExten* next;
  XExtCodes codes;
  void* hooks[9];
  char* name;
  void* more[16];
} XExten;
XExtCodes* XAddExtension(Display* d) {
  XExten* e = (XExten*)calloc(1, sizeof(XExten));
  XExten** head = (XExten**)((char*)d + 0x140);
  static int next_ext = 1;
  e->codes.extension = next_ext++;
  e->codes.major_opcode = 128 + e->codes.extension;
  e->codes.first_event = 64;
  e->codes.first_error = 128;
  e->next = *head;
  *head = e;
  return &e->codes;
}
Bool XQueryExtension(Display* d, const char* name, int* major, int* ev, int* err) {
  (void)d;
  (void)name;
  if (major) *major = 0;
  if (ev) *ev = 0;
  if (err) *err = 0;
  return 0; /* no extensions: in particular no MIT-SHM */
}
Colormap XCreateColormap(Display* d, Window w, Visual* v, int alloc) {
  (void)d; (void)w; (void)v; (void)alloc;
  return g_next_id++;
}
GC XCreateGC(Display* d, Drawable dr, unsigned long mask, void* values) {
  (void)d; (void)dr; (void)mask; (void)values;
  return (GC)calloc(1, 256);
}
int XFreeGC(Display* d, GC gc) {
  (void)d;
  free(gc);
  return 1;
}
static int destroy_image(XImage* im) {
  if (im) {
    free(im->data);
    free(im);
  }
  return 1;
}
static unsigned long get_pixel(XImage* im, int x, int y) {
  if (!im->data) return 0;
  return *(uint32_t*)(im->data + (size_t)y * im->bytes_per_line + (size_t)x * 4);
}
static int put_pixel(XImage* im, int x, int y, unsigned long p) {
  if (im->data) *(uint32_t*)(im->data + (size_t)y * im->bytes_per_line + (size_t)x * 4) = (uint32_t)p;
  return 1;
}
XImage* XCreateImage(Display* d, Visual* v, unsigned depth, int format, int offset, char* data, unsigned width, unsigned height, int pad, int bpl) {
  (void)d;
  XImage* im = (XImage*)calloc(1, sizeof(XImage));
  im->width = (int)width;
  im->height = (int)height;
  im->xoffset = offset;
  im->format = format;
  im->data = data;
  im->byte_order = LSBFirst;
  im->bitmap_unit = 32;
  im->bitmap_bit_order = LSBFirst;
  im->bitmap_pad = pad ? pad : 32;
  im->depth = (int)depth;
  im->bits_per_pixel = 32;
  im->bytes_per_line = bpl ? bpl : (int)width * 4;
  if (v) {
    im->red_mask = v->red_mask;
    im->green_mask = v->green_mask;
    im->blue_mask = v->blue_mask;
  }
  im->f.destroy_image = destroy_image;
  im->f.get_pixel = get_pixel;
  im->f.put_pixel = put_pixel;
  return im;
}
XImage* XShmCreateImage(Display* d, Visual* v, unsigned depth, int format, char* data, void* shminfo, unsigned w, unsigned h) {
  (void)shminfo;
  return XCreateImage(d, v, depth, format, 0, data, w, h, 32, 0);
}
Bool XShmAttach(Display* d, void* info) { (void)d; (void)info; return 0; }
Bool XShmPutImage(Display* d, Drawable dr, GC gc, XImage* im, int sx, int sy, int dx, int dy, unsigned w, unsigned h, Bool ev) {
  (void)d; (void)dr; (void)gc; (void)im; (void)sx; (void)sy; (void)dx; (void)dy; (void)w; (void)h; (void)ev;
  return 1;
}
Pixmap XCreatePixmap(Display* d, Drawable dr, unsigned w, unsigned h, unsigned depth) {
  (void)d; (void)dr; (void)depth;
  return fake_x11_window((int)w, (int)h);
}
int XFreePixmap(Display* d, Pixmap p) { (void)d; (void)p; return 1; }
int XDrawString16(Display* d, Drawable dr, GC gc, int x, int y, const void* s, int n) { (void)d; (void)dr; (void)gc; (void)x; (void)y; (void)s; (void)n; return 1; }
int XFillRectangle(Display* d, Drawable dr, GC gc, int x, int y, unsigned w, unsigned h) { (void)d; (void)dr; (void)gc; (void)x; (void)y; (void)w; (void)h; return 1; }
int XFlush(Display* d) { (void)d; return 1; }
int XSync(Display* d, Bool discard) { (void)d; (void)discard; return 1; }
void* XSynchronize(Display* d, Bool onoff) { (void)d; (void)onoff; return NULL; }
int XFree(void* p) {
  free(p);
  return 1;
}
int XFreeFontInfo(char** names, void* info, int n) { (void)names; (void)info; (void)n; return 1; }
void* XQueryFont(Display* d, XID id) { (void)d; (void)id; return NULL; }
int XGetGeometry(Display* d, Drawable dr, Window* root, int* x, int* y, unsigned* w, unsigned* h, unsigned* border, unsigned* depth) {
  (void)d;
  int ww, hh;
  draw_size(dr, &ww, &hh);
  if (root) *root = 0x100;
  if (x) *x = 0;
  if (y) *y = 0;
  if (w) *w = (unsigned)ww;
  if (h) *h = (unsigned)hh;
  if (border) *border = 0;
  if (depth) *depth = 24;
  return 1;
}
XImage* XGetImage(Display* d, Drawable dr, int x, int y, unsigned w, unsigned h, unsigned long mask, int format) {
  (void)dr; (void)x; (void)y; (void)mask;
  return XCreateImage(d, &g_visual, 24, format, 0, (char*)calloc((size_t)w * h, 4), w, h, 32, 0);
}
XVisualInfo* XGetVisualInfo(Display* d, long mask, XVisualInfo* tmpl, int* n) {
  (void)d;
  /* VisualIDMask 1, VisualScreenMask 2, VisualDepthMask 4, VisualClassMask 8 */
  if (tmpl) {
    if ((mask & 0x1) && tmpl->visualid != g_visual.visualid) goto none;
    if ((mask & 0x2) && tmpl->screen != 0) goto none;
    if ((mask & 0x4) && tmpl->depth != 24) goto none;
    if ((mask & 0x8) && tmpl->c_class != TrueColor) goto none;
  }
  {
    XVisualInfo* vi = (XVisualInfo*)calloc(1, sizeof(XVisualInfo));
    vi->visual = &g_visual;
    vi->visualid = g_visual.visualid;
    vi->screen = 0;
    vi->depth = 24;
    vi->c_class = TrueColor;
    vi->red_mask = g_visual.red_mask;
    vi->green_mask = g_visual.green_mask;
    vi->blue_mask = g_visual.blue_mask;
    vi->colormap_size = 256;
    vi->bits_per_rgb = 8;
    if (n) *n = 1;
    return vi;
  }
none:
  if (n) *n = 0;
  return NULL;
}
int XGetWindowAttributes(Display* d, Window w, XWindowAttributes* a) {
  (void)d;
  memset(a, 0, sizeof(*a));
  draw_size(w, &a->width, &a->height);
  a->depth = 24;
  a->visual = &g_visual;
  a->root = 0x100;
  a->c_class = 1; /* InputOutput */
  a->colormap = 0x20;
  a->map_state = 2; /* IsViewable */
  a->screen = &g_screen;
  return 1;
}
int XPutImage(Display* d, Drawable dr, GC gc, XImage* im, int sx, int sy, int dx, int dy, unsigned w, unsigned h) {
  (void)d; (void)dr; (void)gc; (void)im; (void)sx; (void)sy; (void)dx; (void)dy; (void)w; (void)h;
  return 1;
}
typedef int (*XErrorHandler)(Display*, void*);
XErrorHandler XSetErrorHandler(XErrorHandler h) { (void)h; return NULL; }
int XSetForeground(Display* d, GC gc, unsigned long fg) { (void)d; (void)gc; (void)fg; return 1; }
int XSetFunction(Display* d, GC gc, int fn) { (void)d; (void)gc; (void)fn; return 1; }
