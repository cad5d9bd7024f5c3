// LD_PRELOAD shim (probe aid): logs which files / device nodes the NVIDIA ICD tries to open while initialising.
#define _GNU_SOURCE
#include <dlfcn.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <errno.h>
#include <sys/stat.h>
static int interesting(const char* p) { return p && (strstr(p, "nvidia") || strstr(p, "/dev/") || strstr(p, "dri") || strstr(p, "vulkan") || strstr(p, "/proc/driver")); }
int open(const char* path, int flags, ...) {
    static int (*real)(const char*, int, ...) = 0; if (!real) real = dlsym(RTLD_NEXT, "open");
    va_list ap; va_start(ap, flags); int mode = va_arg(ap, int); va_end(ap);
    int r = real(path, flags, mode); int e = errno;
    if (interesting(path)) fprintf(stderr, "[open] %s -> %d (%s)\n", path, r, r < 0 ? strerror(e) : "ok");
    errno = e; return r;
}
int open64(const char* path, int flags, ...) {
    static int (*real)(const char*, int, ...) = 0; if (!real) real = dlsym(RTLD_NEXT, "open64");
    va_list ap; va_start(ap, flags); int mode = va_arg(ap, int); va_end(ap);
    int r = real(path, flags, mode); int e = errno;
    if (interesting(path)) fprintf(stderr, "[open64] %s -> %d (%s)\n", path, r, r < 0 ? strerror(e) : "ok");
    errno = e; return r;
}
int openat(int dirfd, const char* path, int flags, ...) {
    static int (*real)(int, const char*, int, ...) = 0; if (!real) real = dlsym(RTLD_NEXT, "openat");
    va_list ap; va_start(ap, flags); int mode = va_arg(ap, int); va_end(ap);
    int r = real(dirfd, path, flags, mode); int e = errno;
    if (interesting(path)) fprintf(stderr, "[openat] %s -> %d (%s)\n", path, r, r < 0 ? strerror(e) : "ok");
    errno = e; return r;
}
int access(const char* path, int mode) {
    static int (*real)(const char*, int) = 0; if (!real) real = dlsym(RTLD_NEXT, "access");
    int r = real(path, mode); int e = errno;
    if (interesting(path)) fprintf(stderr, "[access] %s -> %d\n", path, r);
    errno = e; return r;
}
int ioctl(int fd, unsigned long req, ...) {
    static int (*real)(int, unsigned long, ...) = 0; if (!real) real = dlsym(RTLD_NEXT, "ioctl");
    va_list ap; va_start(ap, req); void* arg = va_arg(ap, void*); va_end(ap);
    int r = real(fd, req, arg); int e = errno;
    if (r < 0) fprintf(stderr, "[ioctl] fd %d req 0x%lx -> %d (%s)\n", fd, req, r, strerror(e));
    errno = e; return r;
}
