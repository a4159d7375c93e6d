/* A plain C consumer of include/nfx.h -- what the Rust `extern "C"` block of INTEGRATION.md binds.
 * Built with gcc against libnfx.so by tests/test_c_abi.py. Without a GPU it checks the schema,
 * key formatting, partition and that nfx_create fails with a status code instead of aborting;
 * with a GPU it also runs one nfx_extract through host buffers. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nfx.h"

#define CHECK(c)                                                           \
    do {                                                                   \
        if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    const int want_gpu = argc > 1 && strcmp(argv[1], "--gpu") == 0;
    CHECK(strstr(nfx_version(), "sm_100a") != NULL);
    CHECK(nfx_feature_count(NFX_FS_ALL) == 418);
    CHECK(nfx_feature_count(NFX_FS_GEOMETRY | NFX_FS_COLOR) == 30);
    CHECK(strcmp(nfx_feature_name(NFX_FS_GEOMETRY | NFX_FS_COLOR, 12), "mean_r") == 0);
    CHECK(strcmp(nfx_feature_name(NFX_FS_GLCM, 223), "information_measure_correlation2_1_-1_254") == 0);
    CHECK(nfx_feature_name(NFX_FS_COLOR, 18) == NULL);
    uint32_t bits = 0;
    CHECK(nfx_parse_feature_set("TeXtUrE", &bits) == NFX_OK && bits == NFX_FS_TEXTURE);
    CHECK(nfx_parse_feature_set("colour", &bits) == NFX_ERR_INVALID);
    CHECK(strstr(nfx_last_error(NULL), "not a valid feature set") != NULL);
    char key[64];
    CHECK(nfx_centroid_key(1024.0f, 33.5f, key, sizeof key) == 9 && strcmp(key, "1024,33.5") == 0);
    int64_t b[5];
    CHECK(nfx_partition(1050, 100, 4, b) == NFX_OK && b[0] == 0 && b[1] == 200 && b[4] == 1050);

    nfx_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.patch_size = 64;
    cfg.batch_size = 100;
    nfx_ctx* ctx = NULL;
    int rc = nfx_create(0, &cfg, &ctx);
    if (!want_gpu) {
        if (rc != NFX_OK) {   /* no device: a status code and a message, never an abort or a CPU fallback */
            CHECK(ctx == NULL && strlen(nfx_last_error(NULL)) > 0);
            printf("c_abi_smoke ok (no GPU: %s)\n", nfx_last_error(NULL));
            return 0;
        }
    }
    CHECK(rc == NFX_OK && ctx != NULL);
    CHECK(nfx_compute(ctx, NFX_FS_COLOR) == NFX_ERR_STATE);   /* nothing staged */

    enum { W = 256, H = 256, N = 3 };
    uint8_t* tile = (uint8_t*)malloc(W * H * 3);
    for (int i = 0; i < W * H * 3; ++i) tile[i] = (uint8_t)((i * 2654435761u) >> 24);
    CHECK(nfx_tile_upload(ctx, tile, W, H, 3 * W, 0, 0) == NFX_OK);
    /* three squares of side 20, 10, 30 as closed rings */
    float xy[N * 5 * 2];
    int64_t off[N + 1] = {0, 5, 10, 15};
    const float cx[N] = {100, 40.5f, 200}, cy[N] = {120, 60.25f, 180}, hs[N] = {10, 5, 15};
    for (int n = 0; n < N; ++n) {
        const float px[5] = {cx[n] - hs[n], cx[n] + hs[n], cx[n] + hs[n], cx[n] - hs[n], cx[n] - hs[n]};
        const float py[5] = {cy[n] - hs[n], cy[n] - hs[n], cy[n] + hs[n], cy[n] + hs[n], cy[n] - hs[n]};
        for (int k = 0; k < 5; ++k) { xy[(n * 5 + k) * 2] = px[k]; xy[(n * 5 + k) * 2 + 1] = py[k]; }
    }
    const uint32_t mask = NFX_FS_GEOMETRY | NFX_FS_COLOR;
    const int F = nfx_feature_count(mask);
    float cent[N * 2], *feat = (float*)malloc(sizeof(float) * N * F);
    CHECK(nfx_extract(ctx, N, xy, off, mask, cent, feat) == NFX_OK);
    for (int n = 0; n < N; ++n) {
        const float area = feat[n * F + 0], perim = feat[n * F + 5];
        CHECK(fabsf(area - 4 * hs[n] * hs[n]) < 1e-3f * area && fabsf(perim - 8 * hs[n]) < 1e-3f * perim);
        CHECK(feat[n * F + 12] > 0.f && feat[n * F + 12] < 1.f);   /* mean_r in (0,1) */
    }
    CHECK(nfx_launch_count(ctx) >= 4);
    CHECK(nfx_destroy(ctx) == NFX_OK);
    free(tile);
    free(feat);
    printf("c_abi_smoke ok (GPU)\n");
    return 0;
}
