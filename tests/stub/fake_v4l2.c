/* TEST INFRASTRUCTURE: an LD_PRELOAD stand-in for a Video4Linux2 capture device (neither the build container nor the GPU boxes
 * have one), so that the streaming sequence of csrc/camera.cpp - open, VIDIOC_QUERYCAP, S_FMT, REQBUFS, QUERYBUF + mmap, QBUF,
 * STREAMON, poll, DQBUF / QBUF per frame, STREAMOFF, munmap, close - runs against something that checks it
 * (tests/test_camera_device.py).  The device is /dev/video<FAKE_V4L2_INDEX>; it enforces the state machine a kernel driver
 * enforces (format before buffers, buffers queried and mapped before they are queued, streaming before a dequeue, a buffer is
 * filled only while it is queued, the application may not requeue what it does not hold) and aborts the process with a message
 * when the application breaks it.  Frames come from files: FAKE_V4L2_FRAMES=<dir> holds 000.bin, 001.bin ... (one payload each:
 * a JPEG, or packed YUYV rows with FAKE_V4L2_STRIDE bytes per line), delivered in order; after the last one the device "is
 * unplugged" (DQBUF fails with ENODEV).
 *   FAKE_V4L2_FORMATS   "MJPG,YUYV" (default), "YUYV" (a camera without Motion-JPEG: S_FMT answers with its own format)
 *   FAKE_V4L2_W / _H    the only frame size the sensor has; S_FMT adjusts every request to it, as drivers do
 *   FAKE_V4L2_BAD_EVERY n: every n-th buffer is handed over with V4L2_BUF_FLAG_ERROR and rubbish in it (USB hiccup)
 * The file descriptor is an eventfd, so poll() on it works unmodified: readable while a filled buffer waits. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <linux/videodev2.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/eventfd.h>
#include <sys/mman.h>
#include <unistd.h>

#define NBUF_MAX 8
enum { B_FREE = 0, B_QUERIED, B_QUEUED, B_APP };   /* B_APP: dequeued, owned by the application */

static struct {
    int fd;                 /* -1: closed */
    int have_fmt, streaming, nbuf, next_frame, delivered;
    uint32_t pixfmt;
    int w, h, stride;
    size_t buflen;
    void* map[NBUF_MAX];
    int state[NBUF_MAX];
    int fifo[NBUF_MAX], nfifo;
} dev = {.fd = -1};

static void die(const char* what) {
    fprintf(stderr, "fake_v4l2: protocol violation: %s\n", what);
    abort();
}
static int env_int(const char* n, int d) { const char* v = getenv(n); return v ? atoi(v) : d; }
static int has_format(const char* fourcc) {
    const char* f = getenv("FAKE_V4L2_FORMATS");
    return strstr(f ? f : "MJPG,YUYV", fourcc) != NULL;
}
static void signal_ready(void) {   /* readable <=> streaming and a queued buffer can be filled with a frame (or with the unplug error) */
    uint64_t v;
    while (read(dev.fd, &v, sizeof v) == sizeof v) {}
    if (dev.streaming && dev.nfifo > 0) { v = 1; if (write(dev.fd, &v, sizeof v) != sizeof v) die("eventfd write"); }
}

static int (*real_open)(const char*, int, ...);
static int (*real_ioctl)(int, unsigned long, ...);
static void* (*real_mmap)(void*, size_t, int, int, int, off_t);
static int (*real_close)(int);
__attribute__((constructor)) static void resolve(void) {
    if (real_open) return;
    real_open = dlsym(RTLD_NEXT, "open");
    real_ioctl = dlsym(RTLD_NEXT, "ioctl");
    real_mmap = dlsym(RTLD_NEXT, "mmap");
    real_close = dlsym(RTLD_NEXT, "close");
}

static int is_device_path(const char* path) {
    char want[64];
    snprintf(want, sizeof want, "/dev/video%d", env_int("FAKE_V4L2_INDEX", 42));
    return path && !strcmp(path, want);
}

static int fake_open(void) {
    if (dev.fd >= 0) { errno = EBUSY; return -1; }
    memset(&dev, 0, sizeof dev);
    dev.fd = eventfd(0, EFD_NONBLOCK);
    dev.w = env_int("FAKE_V4L2_W", 64); dev.h = env_int("FAKE_V4L2_H", 48);
    return dev.fd;
}

int open(const char* path, int flags, ...) {
    resolve();
    mode_t mode = 0;
    if (flags & (O_CREAT | O_TMPFILE)) { va_list ap; va_start(ap, flags); mode = va_arg(ap, mode_t); va_end(ap); }
    if (is_device_path(path)) return fake_open();
    return real_open(path, flags, mode);
}
int open64(const char* path, int flags, ...) {
    resolve();
    mode_t mode = 0;
    if (flags & (O_CREAT | O_TMPFILE)) { va_list ap; va_start(ap, flags); mode = va_arg(ap, mode_t); va_end(ap); }
    if (is_device_path(path)) return fake_open();
    return real_open(path, flags | O_LARGEFILE, mode);
}

/* the fortified entry points g++ -O2 -D_FORTIFY_SOURCE routes open(path, flags) to */
int __open_2(const char* path, int flags) { resolve(); return is_device_path(path) ? fake_open() : real_open(path, flags, 0); }
int __open64_2(const char* path, int flags) { resolve(); return is_device_path(path) ? fake_open() : real_open(path, flags | O_LARGEFILE, 0); }

static int load_frame(int idx, void* dst, size_t cap, uint32_t* used) {
    char path[1024];
    snprintf(path, sizeof path, "%s/%03d.bin", getenv("FAKE_V4L2_FRAMES") ? getenv("FAKE_V4L2_FRAMES") : ".", idx);
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    const size_t n = fread(dst, 1, cap, f);
    fclose(f);
    *used = (uint32_t)n;
    return 0;
}

static int fake_ioctl(unsigned long req, void* arg) {
    switch (req) {
    case VIDIOC_QUERYCAP: {
        struct v4l2_capability* c = arg;
        memset(c, 0, sizeof *c);
        strcpy((char*)c->driver, "fake_v4l2"); strcpy((char*)c->card, "test sensor");
        c->capabilities = V4L2_CAP_VIDEO_CAPTURE | V4L2_CAP_STREAMING | V4L2_CAP_DEVICE_CAPS;
        c->device_caps = V4L2_CAP_VIDEO_CAPTURE | V4L2_CAP_STREAMING;
        return 0;
    }
    case VIDIOC_S_FMT: {
        struct v4l2_format* f = arg;
        if (f->type != V4L2_BUF_TYPE_VIDEO_CAPTURE) { errno = EINVAL; return -1; }
        if (dev.nbuf) { errno = EBUSY; return -1; }   /* the format is fixed once buffers exist */
        uint32_t pf = f->fmt.pix.pixelformat;
        if (!((pf == V4L2_PIX_FMT_MJPEG && has_format("MJPG")) || (pf == V4L2_PIX_FMT_YUYV && has_format("YUYV"))))   /* drivers answer with what they have */
            pf = has_format("MJPG") ? V4L2_PIX_FMT_MJPEG : (has_format("YUYV") ? V4L2_PIX_FMT_YUYV : V4L2_PIX_FMT_NV12);
        dev.pixfmt = pf;
        dev.stride = pf == V4L2_PIX_FMT_YUYV ? env_int("FAKE_V4L2_STRIDE", 2 * dev.w) : 0;
        dev.buflen = pf == V4L2_PIX_FMT_YUYV ? (size_t)dev.stride * dev.h : (size_t)dev.w * dev.h * 2;
        f->fmt.pix.width = dev.w; f->fmt.pix.height = dev.h; f->fmt.pix.pixelformat = pf; f->fmt.pix.field = V4L2_FIELD_NONE;
        f->fmt.pix.bytesperline = dev.stride; f->fmt.pix.sizeimage = (uint32_t)dev.buflen;
        dev.have_fmt = 1;
        return 0;
    }
    case VIDIOC_REQBUFS: {
        struct v4l2_requestbuffers* r = arg;
        if (!dev.have_fmt) die("VIDIOC_REQBUFS before VIDIOC_S_FMT");
        if (r->type != V4L2_BUF_TYPE_VIDEO_CAPTURE || r->memory != V4L2_MEMORY_MMAP) { errno = EINVAL; return -1; }
        if (dev.streaming) { errno = EBUSY; return -1; }
        if (r->count > 3) r->count = 3;                 /* drivers grant what they like */
        dev.nbuf = (int)r->count;
        for (int i = 0; i < dev.nbuf; i++) dev.state[i] = B_FREE;
        return 0;
    }
    case VIDIOC_QUERYBUF: {
        struct v4l2_buffer* b = arg;
        if (b->type != V4L2_BUF_TYPE_VIDEO_CAPTURE || (int)b->index >= dev.nbuf) { errno = EINVAL; return -1; }
        b->memory = V4L2_MEMORY_MMAP; b->length = (uint32_t)dev.buflen; b->m.offset = b->index * 0x100000u;
        if (dev.state[b->index] == B_FREE) dev.state[b->index] = B_QUERIED;
        return 0;
    }
    case VIDIOC_QBUF: {
        struct v4l2_buffer* b = arg;
        if (b->type != V4L2_BUF_TYPE_VIDEO_CAPTURE || b->memory != V4L2_MEMORY_MMAP || (int)b->index >= dev.nbuf) { errno = EINVAL; return -1; }
        if (dev.state[b->index] == B_FREE) die("VIDIOC_QBUF of a buffer that was never queried");
        if (!dev.map[b->index]) die("VIDIOC_QBUF of a buffer that is not mapped");
        if (dev.state[b->index] == B_QUEUED) die("VIDIOC_QBUF of a buffer the driver already holds");
        dev.state[b->index] = B_QUEUED;
        dev.fifo[dev.nfifo++] = (int)b->index;
        signal_ready();
        return 0;
    }
    case VIDIOC_STREAMON:
        if (!dev.nbuf) die("VIDIOC_STREAMON without buffers");
        dev.streaming = 1;
        signal_ready();
        return 0;
    case VIDIOC_STREAMOFF:
        dev.streaming = 0;
        for (int i = 0; i < dev.nbuf; i++) if (dev.state[i] == B_QUEUED || dev.state[i] == B_APP) dev.state[i] = B_QUERIED;
        dev.nfifo = 0;
        signal_ready();
        return 0;
    case VIDIOC_DQBUF: {
        struct v4l2_buffer* b = arg;
        if (!dev.streaming) die("VIDIOC_DQBUF while not streaming");
        if (b->type != V4L2_BUF_TYPE_VIDEO_CAPTURE || b->memory != V4L2_MEMORY_MMAP) { errno = EINVAL; return -1; }
        if (dev.nfifo == 0) { errno = EAGAIN; return -1; }   /* O_NONBLOCK: nothing filled */
        const int idx = dev.fifo[0];
        uint32_t used = 0;
        const int bad_every = env_int("FAKE_V4L2_BAD_EVERY", 0);
        const int bad = bad_every > 0 && (dev.delivered + 1) % bad_every == 0;
        if (bad) {
            memset(dev.map[idx], 0xA5, dev.buflen);
            used = (uint32_t)dev.buflen / 3;
        } else if (load_frame(dev.next_frame, dev.map[idx], dev.buflen, &used)) {
            errno = ENODEV;                                   /* the camera was unplugged */
            return -1;
        } else {
            dev.next_frame++;
        }
        dev.delivered++;
        memmove(dev.fifo, dev.fifo + 1, sizeof(int) * (size_t)(--dev.nfifo));
        dev.state[idx] = B_APP;
        memset(b, 0, sizeof *b);
        b->type = V4L2_BUF_TYPE_VIDEO_CAPTURE; b->memory = V4L2_MEMORY_MMAP; b->index = (uint32_t)idx; b->bytesused = used;
        b->length = (uint32_t)dev.buflen; b->flags = V4L2_BUF_FLAG_MAPPED | (bad ? V4L2_BUF_FLAG_ERROR : 0); b->sequence = (uint32_t)dev.delivered;
        signal_ready();
        return 0;
    }
    default:
        errno = ENOTTY;
        return -1;
    }
}

int ioctl(int fd, unsigned long req, ...) {
    resolve();
    va_list ap;
    va_start(ap, req);
    void* arg = va_arg(ap, void*);
    va_end(ap);
    if (dev.fd >= 0 && fd == dev.fd) return fake_ioctl(req, arg);
    return real_ioctl(fd, req, arg);
}

void* mmap(void* addr, size_t len, int prot, int flags, int fd, off_t off) {
    resolve();
    if (dev.fd >= 0 && fd == dev.fd) {
        const int idx = (int)(off / 0x100000);
        if (idx < 0 || idx >= dev.nbuf || off % 0x100000 || dev.state[idx] == B_FREE) die("mmap of an offset VIDIOC_QUERYBUF never returned");
        if (len != dev.buflen) die("mmap length differs from v4l2_buffer.length");
        if (!(flags & MAP_SHARED)) die("capture buffers must be mapped MAP_SHARED");
        void* p = real_mmap(NULL, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) dev.map[idx] = p;
        return p;
    }
    return real_mmap(addr, len, prot, flags, fd, off);
}

int close(int fd) {
    resolve();
    if (dev.fd >= 0 && fd == dev.fd) {
        if (dev.streaming) fprintf(stderr, "fake_v4l2: closed while streaming (allowed, but VIDIOC_STREAMOFF was expected)\n");
        fprintf(stderr, "fake_v4l2: closed after %d frames\n", dev.delivered);
        dev.fd = -1;
    }
    return real_close(fd);
}
