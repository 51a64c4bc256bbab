/* test_c_api_rtx.c — SURVEY.md §8 f3: the reference's FLAT C API (include/luisa/api/api.h, src/api/runtime.cpp:487-568:
 * luisa_compute_mesh_create / accel_create / stream_dispatch with LCMeshBuildCommand / LCAccelBuildCommand) driving the
 * `b200rt` backend plugin from plain C.  liblc-api.so is the reference's own src/api/runtime.cpp + logging.cpp compiled
 * where they lie (plugin/Makefile); nothing of it is copied here.
 *
 * Scene and rays: test_rtx.cpp:24-28, 101-107 (one triangle, two instances) with W x H orthographic rays (:67-73, jitter
 * replaced by +0.5).  The program builds the mesh and the accel through the C API's command list, traces through the
 * plugin's flat entry point b200rt_trace (the C API has no custom-command variant), downloads the 24-byte hit records with
 * an LCBufferDownloadCommand and writes them to <out>; tests/test_gpu_plugin.py compares them with the CPU oracle.
 * usage: test_c_api_rtx <out.bin> [W H]
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <luisa/api/api.h>

typedef int (*b200rt_trace_fn)(void *device, uint64_t stream, uint32_t mode, uint64_t accel, uint64_t ray_buffer, uint64_t hit_buffer,
                               uint64_t count, uint32_t visibility_mask);

static void dispatch(LCDevice dev, LCStream stream, const LCCommand *cmds, size_t n) {
    LCCommandList list;
    list.commands = cmds, list.commands_count = n;
    luisa_compute_stream_dispatch(dev, stream, list, NULL, NULL);
}

static void affine_identity(float *a) {
    memset(a, 0, 12 * sizeof(float));
    a[0] = a[5] = a[10] = 1.0f;
}

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <out.bin> [W H]\n", argv[0]);
        return 2;
    }
    const uint32_t W = argc > 3 ? (uint32_t)atoi(argv[2]) : 256u, H = argc > 3 ? (uint32_t)atoi(argv[3]) : 256u;
    const uint64_t n = (uint64_t)W * H;
    LCContext ctx = luisa_compute_context_create(argv[0]);
    LCDevice dev = luisa_compute_device_create(ctx, "b200rt", "{}");
    LCStream stream = {luisa_compute_stream_create(dev, LC_STREAM_TAG_COMPUTE).handle};

    /* geometry: float3 at stride 16, one triangle */
    const float verts[12] = {-0.5f, -0.5f, 0.0f, 0.0f, 0.5f, -0.5f, 0.0f, 0.0f, 0.0f, 0.5f, 0.0f, 0.0f};
    const uint32_t tri[3] = {0u, 1u, 2u};
    float *rays = (float *)malloc(n * 32u);
    for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) {
            float *r = rays + 8u * ((uint64_t)y * W + x);
            r[0] = ((float)x + 0.5f) / (float)W * 2.0f - 1.0f, r[1] = -(((float)y + 0.5f) / (float)H * 2.0f - 1.0f), r[2] = 1.0f, r[3] = 0.0f;
            r[4] = 0.0f, r[5] = 0.0f, r[6] = -1.0f, r[7] = 3.402823466e38f;
        }
    LCBuffer vbuf = {luisa_compute_buffer_create(dev, NULL, sizeof(verts), NULL).resource.handle};
    LCBuffer tbuf = {luisa_compute_buffer_create(dev, NULL, sizeof(tri), NULL).resource.handle};
    LCBuffer rbuf = {luisa_compute_buffer_create(dev, NULL, n * 32u, NULL).resource.handle};
    LCBuffer hbuf = {luisa_compute_buffer_create(dev, NULL, n * 24u, NULL).resource.handle};
    LCBuffer obuf = {luisa_compute_buffer_create(dev, NULL, n * 4u, NULL).resource.handle};

    LCAccelOption option;
    option.hint = LC_ACCEL_USAGE_HINT_FAST_TRACE, option.allow_compaction = true, option.allow_update = false;
    LCMesh mesh = {luisa_compute_mesh_create(dev, &option).handle};
    LCAccel accel = {luisa_compute_accel_create(dev, &option).handle};

    /* instances of test_rtx.cpp:103-105: scaling(1.5); translation(-0.25, 0, 0.1) * rotation(z, 0.5) */
    LCAccelBuildModification mods[2];
    memset(mods, 0, sizeof(mods));
    for (int i = 0; i < 2; i++) {
        mods[i].index = (uint32_t)i, mods[i].user_id = 0u, mods[i].visibility = 0xffu, mods[i].mesh = mesh._0;
        mods[i].flags.bits = (1u << 0) | (1u << 1) | (1u << 2) | (1u << 4) | (1u << 5); /* primitive, transform, opaque_on, visibility, user_id */
        affine_identity(mods[i].affine);
    }
    mods[0].affine[0] = mods[0].affine[5] = mods[0].affine[10] = 1.5f;
    {
        const float c = (float)cos(0.5), s = (float)sin(0.5);
        float *a = mods[1].affine; /* row-major 3x4 */
        a[0] = c, a[1] = -s, a[3] = -0.25f;
        a[4] = s, a[5] = c, a[7] = 0.0f;
        a[10] = 1.0f, a[11] = 0.1f;
    }

    LCCommand cmds[5];
    memset(cmds, 0, sizeof(cmds));
    cmds[0].tag = LC_COMMAND_BUFFER_UPLOAD;
    cmds[0].buffer_upload.buffer = vbuf, cmds[0].buffer_upload.offset = 0, cmds[0].buffer_upload.size = sizeof(verts), cmds[0].buffer_upload.data = (const uint8_t *)verts;
    cmds[1].tag = LC_COMMAND_BUFFER_UPLOAD;
    cmds[1].buffer_upload.buffer = tbuf, cmds[1].buffer_upload.offset = 0, cmds[1].buffer_upload.size = sizeof(tri), cmds[1].buffer_upload.data = (const uint8_t *)tri;
    cmds[2].tag = LC_COMMAND_BUFFER_UPLOAD;
    cmds[2].buffer_upload.buffer = rbuf, cmds[2].buffer_upload.offset = 0, cmds[2].buffer_upload.size = n * 32u, cmds[2].buffer_upload.data = (const uint8_t *)rays;
    cmds[3].tag = LC_COMMAND_MESH_BUILD;
    cmds[3].mesh_build.mesh = mesh, cmds[3].mesh_build.request = LC_ACCEL_BUILD_REQUEST_PREFER_UPDATE;
    cmds[3].mesh_build.vertex_buffer = vbuf, cmds[3].mesh_build.vertex_buffer_offset = 0, cmds[3].mesh_build.vertex_buffer_size = sizeof(verts);
    cmds[3].mesh_build.vertex_stride = 16;
    cmds[3].mesh_build.index_buffer = tbuf, cmds[3].mesh_build.index_buffer_offset = 0, cmds[3].mesh_build.index_buffer_size = sizeof(tri);
    cmds[3].mesh_build.index_stride = 12;
    cmds[4].tag = LC_COMMAND_ACCEL_BUILD;
    cmds[4].accel_build.accel = accel, cmds[4].accel_build.request = LC_ACCEL_BUILD_REQUEST_PREFER_UPDATE, cmds[4].accel_build.instance_count = 2u;
    cmds[4].accel_build.modifications = mods, cmds[4].accel_build.modifications_count = 2, cmds[4].accel_build.update_instance_buffer_only = false;
    dispatch(dev, stream, cmds, 5);

    /* the plugin is loaded by now: its flat trace entry */
    void *plugin = dlopen("liblc-backend-b200rt.so", RTLD_NOW | RTLD_NOLOAD);
    if (!plugin) plugin = dlopen("liblc-backend-b200rt.so", RTLD_NOW);
    b200rt_trace_fn trace = plugin ? (b200rt_trace_fn)dlsym(plugin, "b200rt_trace") : NULL;
    if (!trace) {
        fprintf(stderr, "b200rt_trace not found: %s\n", dlerror());
        return 3;
    }
    if (trace((void *)dev._0, stream._0, 0u, accel._0, rbuf._0, hbuf._0, n, 0xffu) != 0) return 4;
    if (trace((void *)dev._0, stream._0, 1u, accel._0, rbuf._0, obuf._0, n, 0xffu) != 0) return 5;

    uint8_t *hits = (uint8_t *)malloc(n * 24u);
    uint32_t *occ = (uint32_t *)malloc(n * 4u);
    LCCommand down[2];
    memset(down, 0, sizeof(down));
    down[0].tag = LC_COMMAND_BUFFER_DOWNLOAD;
    down[0].buffer_download.buffer = hbuf, down[0].buffer_download.offset = 0, down[0].buffer_download.size = n * 24u, down[0].buffer_download.data = hits;
    down[1].tag = LC_COMMAND_BUFFER_DOWNLOAD;
    down[1].buffer_download.buffer = obuf, down[1].buffer_download.offset = 0, down[1].buffer_download.size = n * 4u, down[1].buffer_download.data = (uint8_t *)occ;
    dispatch(dev, stream, down, 2);
    luisa_compute_stream_synchronize(dev, stream);

    uint64_t nhit = 0, nocc = 0;
    for (uint64_t i = 0; i < n; i++) {
        nhit += *(const uint32_t *)(hits + 24u * i) != 0xffffffffu;
        nocc += occ[i] != 0u;
    }
    FILE *f = fopen(argv[1], "wb");
    if (!f) return 6;
    /* everything the checker needs to rebuild the same case bit for bit: size, instance transforms, rays, then the results */
    fwrite(&W, 4, 1, f), fwrite(&H, 4, 1, f);
    fwrite(mods[0].affine, 4, 12, f), fwrite(mods[1].affine, 4, 12, f);
    fwrite(rays, 32, n, f);
    fwrite(hits, 24, n, f), fwrite(occ, 4, n, f);
    fclose(f);
    printf("c-api: %llu rays, %llu hits, %llu occluded\n", (unsigned long long)n, (unsigned long long)nhit, (unsigned long long)nocc);

    luisa_compute_accel_destroy(dev, accel);
    luisa_compute_mesh_destroy(dev, mesh);
    luisa_compute_buffer_destroy(dev, vbuf), luisa_compute_buffer_destroy(dev, tbuf), luisa_compute_buffer_destroy(dev, rbuf);
    luisa_compute_buffer_destroy(dev, hbuf), luisa_compute_buffer_destroy(dev, obuf);
    luisa_compute_stream_destroy(dev, stream);
    luisa_compute_device_destroy(dev);
    luisa_compute_context_destroy(ctx);
    free(rays), free(hits), free(occ);
    return nhit > 0 && nhit == nocc ? 0 : 7;
}
