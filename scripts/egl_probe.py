"""Probe: can a headless OpenGL context be had on this box from NVIDIA's EGL vendor library alone?

The image has no libglvnd (no libEGL.so.1 / libGL.so.1), but the driver's vendor library libEGL_nvidia.so.0 is there.  It
exports only the glvnd vendor entry point __egl_Main; this script plays the part of libglvnd's EGL front end (the few
call-backs a vendor expects) and asks the vendor for its own eglXxx / glXxx entry points.  Prints what works; exits 0 when a
current OpenGL context exists."""
import ctypes as C
import sys

PATHS = ("/usr/lib/libEGL_nvidia.so.0", "/usr/local/nvidia/lib64/libEGL_nvidia.so.0", "/usr/lib/x86_64-linux-gnu/libEGL_nvidia.so.0")
EGL_PLATFORM_DEVICE_EXT = 0x313F
EGL_OPENGL_API = 0x30A2
EGL_NONE = 0x3038
EGL_SURFACE_TYPE, EGL_PBUFFER_BIT = 0x3033, 0x0001
EGL_RENDERABLE_TYPE, EGL_OPENGL_BIT = 0x3040, 0x0008
EGL_CONTEXT_MAJOR_VERSION, EGL_CONTEXT_MINOR_VERSION = 0x3098, 0x30FB
EGL_EXTENSIONS, EGL_VENDOR, EGL_VERSION = 0x3055, 0x3053, 0x3054


class State:
    api = EGL_OPENGL_API
    ctx = None
    dpy = None
    err = 0x3000
    keep = []


def load():
    for p in PATHS:
        try:
            return C.CDLL(p, mode=C.RTLD_GLOBAL)
        except OSError:
            continue
    return None


def main():
    lib = load()
    if lib is None:
        print("no libEGL_nvidia.so.0"); return 2
    vp = C.c_void_p
    vendor_handle = C.create_string_buffer(64)          # opaque __EGLvendorInfo for the vendor to hand back

    def cb(restype, *argtypes):
        def deco(fn):
            f = C.CFUNCTYPE(restype, *argtypes)(fn)
            State.keep.append(f)
            return f
        return deco
    exports = (vp * 24)()
    fns = [
        cb(None)(lambda: None),                                                # threadInit
        cb(C.c_uint)(lambda: State.api),                                       # getCurrentApi
        cb(vp)(lambda: C.addressof(vendor_handle) if State.ctx else None),     # getCurrentVendor
        cb(vp)(lambda: State.ctx),                                             # getCurrentContext
        cb(vp)(lambda: State.dpy),                                             # getCurrentDisplay
        cb(vp, C.c_int)(lambda rd: None),                                      # getCurrentSurface
        cb(vp, vp, C.c_int)(lambda v, i: None),                                # fetchDispatchEntry
        cb(C.c_uint, C.c_int)(lambda e: (setattr(State, "err", e), 1)[1]),     # setEGLError
        cb(C.c_uint, vp)(lambda v: 1),                                         # setLastVendor
        cb(C.c_uint, vp, vp)(lambda d, v: 1),                                  # setVendorForDisplay
        cb(C.c_uint, vp, vp)(lambda d, v: 1),                                  # setVendorForDevice
        cb(vp, vp)(lambda d: C.addressof(vendor_handle)),                      # getVendorFromDisplay
        cb(vp, vp)(lambda d: C.addressof(vendor_handle)),                      # getVendorFromDevice
    ]
    for i, f in enumerate(fns):
        exports[i] = C.cast(f, vp)
    imports = (vp * 32)()
    lib.__egl_Main.restype = C.c_uint
    lib.__egl_Main.argtypes = [C.c_uint32, vp, vp, vp]
    ok = 0
    for version in (1, 0, 2, 3):                         # (major 0 << 16) | minor
        ok = lib.__egl_Main(version, C.cast(exports, vp), C.cast(vendor_handle, vp), C.cast(imports, vp))
        print("__egl_Main(version %d) ->" % version, ok, flush=True)
        if ok:
            break
    if not ok:
        return 3
    print("imports:", [hex(x or 0) for x in imports[:12]], flush=True)
    get_vendor_string = C.CFUNCTYPE(C.c_char_p, C.c_int)(imports[2])
    print("platform extensions:", get_vendor_string(0), flush=True)
    gpa = C.CFUNCTYPE(vp, C.c_char_p)(imports[3])

    def egl(name, restype, *argtypes):
        a = gpa(name.encode())
        if not a:
            raise RuntimeError(name + " not found")
        return C.CFUNCTYPE(restype, *argtypes)(a)
    eglQueryDevicesEXT = egl("eglQueryDevicesEXT", C.c_uint, C.c_int, C.POINTER(vp), C.POINTER(C.c_int))
    devs = (vp * 16)()
    n = C.c_int(0)
    print("eglQueryDevicesEXT ->", eglQueryDevicesEXT(16, devs, C.byref(n)), "devices", n.value, "err", hex(State.err), flush=True)
    get_platform_display = C.CFUNCTYPE(vp, C.c_uint, vp, vp)(imports[0])
    eglInitialize = egl("eglInitialize", C.c_uint, vp, C.POINTER(C.c_int), C.POINTER(C.c_int))
    eglQueryString = egl("eglQueryString", C.c_char_p, vp, C.c_int)
    dpy = None
    if n.value < 1:
        for plat, name in ((0x31DD, "surfaceless"), (EGL_PLATFORM_DEVICE_EXT, "device/default")):
            d = get_platform_display(plat, None, None)
            ma, mi = C.c_int(0), C.c_int(0)
            r = eglInitialize(d, C.byref(ma), C.byref(mi)) if d else 0
            print("platform", name, "display", d, "eglInitialize ->", r, "err", hex(State.err), flush=True)
            if r:
                dpy = d
                break
        if not dpy:
            return 4
    for i in range(n.value if not dpy else 0):
        d = get_platform_display(EGL_PLATFORM_DEVICE_EXT, devs[i], None)
        ma, mi = C.c_int(0), C.c_int(0)
        r = eglInitialize(d, C.byref(ma), C.byref(mi)) if d else 0
        print("device", i, "display", d, "eglInitialize ->", r, "EGL %d.%d" % (ma.value, mi.value), "err", hex(State.err), flush=True)
        if r:
            dpy = d
            break
    if not dpy:
        return 5
    State.dpy = dpy
    print("EGL vendor:", eglQueryString(dpy, EGL_VENDOR), "version:", eglQueryString(dpy, EGL_VERSION), flush=True)
    eglBindAPI = egl("eglBindAPI", C.c_uint, C.c_uint)
    print("eglBindAPI(OPENGL) ->", eglBindAPI(EGL_OPENGL_API), flush=True)
    eglChooseConfig = egl("eglChooseConfig", C.c_uint, vp, C.POINTER(C.c_int), C.POINTER(vp), C.c_int, C.POINTER(C.c_int))
    attrs = (C.c_int * 5)(EGL_SURFACE_TYPE, EGL_PBUFFER_BIT, EGL_RENDERABLE_TYPE, EGL_OPENGL_BIT, EGL_NONE)
    cfg = (vp * 1)()
    nc = C.c_int(0)
    print("eglChooseConfig ->", eglChooseConfig(dpy, attrs, cfg, 1, C.byref(nc)), nc.value, flush=True)
    eglCreateContext = egl("eglCreateContext", vp, vp, vp, vp, C.POINTER(C.c_int))
    cattrs = (C.c_int * 5)(EGL_CONTEXT_MAJOR_VERSION, 4, EGL_CONTEXT_MINOR_VERSION, 3, EGL_NONE)
    ctx = eglCreateContext(dpy, cfg[0] if nc.value else None, None, cattrs)
    print("eglCreateContext ->", ctx, "err", hex(State.err), flush=True)
    if not ctx:
        return 6
    eglMakeCurrent = egl("eglMakeCurrent", C.c_uint, vp, vp, vp, vp)
    r = eglMakeCurrent(dpy, None, None, ctx)
    print("eglMakeCurrent(surfaceless) ->", r, "err", hex(State.err), flush=True)
    if not r:
        return 7
    State.ctx = ctx
    glGetString = C.CFUNCTYPE(C.c_char_p, C.c_uint)(gpa(b"glGetString"))
    print("GL_VENDOR:", glGetString(0x1F00), "GL_RENDERER:", glGetString(0x1F01), "GL_VERSION:", glGetString(0x1F02), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
