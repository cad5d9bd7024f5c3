// Loader-less probe of the NVIDIA Vulkan ICD on the GPU box: dlopen the driver's ICD library
// directly (no libvulkan, no icd.json, no Vulkan headers in the image: the few prototypes used are declared here)
// and ask it for its physical devices. Build: gcc -O1 -o oracle/vk/vk_probe oracle/vk/vk_probe.c -ldl
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef void (*PFN_vkVoidFunction)(void);
typedef PFN_vkVoidFunction (*PFN_GetInstanceProcAddr)(void* instance, const char* name);
typedef int32_t (*PFN_Negotiate)(uint32_t* version);

typedef struct { uint32_t sType; const void* pNext; const char* pApplicationName; uint32_t applicationVersion;
                 const char* pEngineName; uint32_t engineVersion; uint32_t apiVersion; } VkApplicationInfo;
typedef struct { uint32_t sType; const void* pNext; uint32_t flags; const VkApplicationInfo* pApplicationInfo;
                 uint32_t enabledLayerCount; const char* const* ppEnabledLayerNames;
                 uint32_t enabledExtensionCount; const char* const* ppEnabledExtensionNames; } VkInstanceCreateInfo;
typedef int32_t (*PFN_vkCreateInstance)(const VkInstanceCreateInfo*, const void*, void** instance);
typedef int32_t (*PFN_vkEnumeratePhysicalDevices)(void* instance, uint32_t* count, void** devices);
typedef void (*PFN_vkGetPhysicalDeviceProperties)(void* phys, void* props);
typedef struct { uint32_t queueFlags, queueCount, timestampValidBits; uint32_t minGran[3]; } VkQueueFamilyProperties;
typedef void (*PFN_vkGetPhysicalDeviceQueueFamilyProperties)(void* phys, uint32_t* count, VkQueueFamilyProperties* props);

int main(void) {
    const char* names[] = {"libGLX_nvidia.so.0", "/usr/lib/libGLX_nvidia.so.0", "/usr/lib/x86_64-linux-gnu/libGLX_nvidia.so.0", "libEGL_nvidia.so.0"};
    void* h = NULL;
    for (unsigned i = 0; i < sizeof(names) / sizeof(names[0]) && !h; ++i) {
        h = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
        printf("dlopen %s: %s\n", names[i], h ? "ok" : dlerror());
    }
    if (!h) return 2;
    PFN_Negotiate nego = (PFN_Negotiate)dlsym(h, "vk_icdNegotiateLoaderICDInterfaceVersion");
    PFN_GetInstanceProcAddr gipa = (PFN_GetInstanceProcAddr)dlsym(h, "vk_icdGetInstanceProcAddr");
    printf("vk_icdNegotiateLoaderICDInterfaceVersion %p, vk_icdGetInstanceProcAddr %p\n", (void*)nego, (void*)gipa);
    if (!gipa) return 3;
    if (nego) { uint32_t v = 5; int32_t r = nego(&v); printf("negotiate -> %d, interface version %u\n", r, v); }
    PFN_vkCreateInstance createInstance = (PFN_vkCreateInstance)gipa(NULL, "vkCreateInstance");
    printf("vkCreateInstance %p\n", (void*)createInstance);
    if (!createInstance) return 4;
    VkApplicationInfo app = {0 /*APPLICATION_INFO*/, NULL, "orbit-oracle", 1, "none", 1, (1u << 22) | (3u << 12)};
    VkInstanceCreateInfo ici = {1 /*INSTANCE_CREATE_INFO*/, NULL, 0, &app, 0, NULL, 0, NULL};
    void* inst = NULL;
    int32_t r = createInstance(&ici, NULL, &inst);
    printf("vkCreateInstance -> %d, instance %p\n", r, inst);
    if (r != 0) return 5;
    PFN_vkEnumeratePhysicalDevices enumerate = (PFN_vkEnumeratePhysicalDevices)gipa(inst, "vkEnumeratePhysicalDevices");
    PFN_vkGetPhysicalDeviceProperties props = (PFN_vkGetPhysicalDeviceProperties)gipa(inst, "vkGetPhysicalDeviceProperties");
    PFN_vkGetPhysicalDeviceQueueFamilyProperties qprops = (PFN_vkGetPhysicalDeviceQueueFamilyProperties)gipa(inst, "vkGetPhysicalDeviceQueueFamilyProperties");
    uint32_t n = 0;
    r = enumerate(inst, &n, NULL);
    printf("vkEnumeratePhysicalDevices -> %d, %u device(s)\n", r, n);
    void* devs[16]; if (n > 16) n = 16;
    r = enumerate(inst, &n, devs);
    for (uint32_t i = 0; i < n; ++i) {
        static uint64_t buf[512];
        memset(buf, 0, sizeof(buf));
        props(devs[i], buf);
        const uint32_t* w = (const uint32_t*)buf;
        printf("device %u: api %u.%u.%u driver 0x%x vendor 0x%x id 0x%x type %u name '%s'\n", i, w[0] >> 22, (w[0] >> 12) & 0x3ff, w[0] & 0xfff,
               w[1], w[2], w[3], w[4], (const char*)(w + 5));
        uint32_t nq = 0; qprops(devs[i], &nq, NULL);
        VkQueueFamilyProperties q[16]; if (nq > 16) nq = 16; qprops(devs[i], &nq, q);
        for (uint32_t k = 0; k < nq; ++k) printf("   queue family %u: flags 0x%x count %u timestampValidBits %u\n", k, q[k].queueFlags, q[k].queueCount, q[k].timestampValidBits);
    }
    return 0;
}
