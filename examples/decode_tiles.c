/* decode_tiles.c — the C ABI used from plain C, no Python and no torch: decode a batch of .covt tiles on one GPU and print,
 * per layer, what CovtParser.decodeCovt (CovtParser.java:53) would have materialised as JTS objects.
 *
 *   gcc -O2 -I include examples/decode_tiles.c -o decode_tiles -L cov-tiles_b200 -lcovt_b200 -Wl,-rpath,$PWD/cov-tiles_b200
 *   ./decode_tiles [--gen3] [--flags N] tile1.covt tile2.covt ...
 *
 * Output: one line per layer "tile layer name features parts rings vertices coords status fnv(coords)". The same binding in
 * Java (Panama FFM) is integration/java/CovtGpuDecoder.java. tests/test_gpu_c_example.py compares the lines with the oracle. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "covt_b200.h"

static void die(covt_ctx* ctx, const char* what, int32_t rc)
{
    char msg[512] = "";
    covt_last_error(ctx, msg, sizeof msg);
    fprintf(stderr, "%s failed: status %d: %s\n", what, rc, msg);
    exit(1);
}

int main(int argc, char** argv)
{
    uint32_t container = COVT_CONTAINER_GEN2B, flags = COVT_FLAG_DEFAULT;
    int first = 1;
    while (first < argc && argv[first][0] == '-') {
        if (!strcmp(argv[first], "--gen3")) container = COVT_CONTAINER_GEN3;
        else if (!strcmp(argv[first], "--flags") && first + 1 < argc) flags = (uint32_t)strtoul(argv[++first], NULL, 0);
        else { fprintf(stderr, "unknown option %s\n", argv[first]); return 2; }
        first++;
    }
    const uint32_t n_tiles = (uint32_t)(argc - first);
    if (!n_tiles) { fprintf(stderr, "usage: %s [--gen3] [--flags N] tile.covt ...\n", argv[0]); return 2; }

    /* tiles back to back + tile_offsets[n + 1] */
    uint64_t* offs = (uint64_t*)calloc(n_tiles + 1, sizeof(uint64_t));
    uint8_t* blob = NULL;
    uint64_t len = 0;
    for (uint32_t i = 0; i < n_tiles; i++) {
        FILE* f = fopen(argv[first + i], "rb");
        if (!f) { perror(argv[first + i]); return 1; }
        fseek(f, 0, SEEK_END);
        const long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        blob = (uint8_t*)realloc(blob, len + (uint64_t)sz + 1);
        if (fread(blob + len, 1, (size_t)sz, f) != (size_t)sz) { perror("fread"); return 1; }
        fclose(f);
        offs[i] = len;
        len += (uint64_t)sz;
    }
    offs[n_tiles] = len;

    covt_ctx* ctx = NULL;
    int32_t rc = covt_create(0, &ctx);
    if (rc != COVT_OK) die(NULL, "covt_create", rc); /* no GPU, no decode: there is no CPU fallback */
    covt_host_register(ctx, blob, len);              /* page-lock the caller's buffer: full PCIe rate */
    covt_result* res = NULL;
    rc = covt_decode_batch(ctx, blob, offs, n_tiles, container, NULL, flags, &res);
    if (rc != COVT_OK) die(ctx, "covt_decode_batch", rc);

    const covt_layer* layers = NULL;
    const uint32_t *tile_status = NULL, *first_layer = NULL;
    if ((rc = covt_result_layers(res, &layers)) != COVT_OK) die(ctx, "covt_result_layers", rc);
    if ((rc = covt_result_tile_status(res, &tile_status, &first_layer)) != COVT_OK) die(ctx, "covt_result_tile_status", rc);
    const uint32_t n_layers = covt_result_num_layers(res);
    for (uint32_t l = 0; l < n_layers; l++) {
        const covt_layer* L = &layers[l];
        /* assembled coordinates of the layer: i32 x,y interleaved, closing vertices included */
        const uint64_t n_ints = 2ull * L->n_coords;
        int32_t* xy = (int32_t*)malloc((n_ints + 1) * sizeof(int32_t));
        if (L->status == COVT_OK && n_ints &&
            (rc = covt_result_read(res, COVT_BUF_A_COORDS, L->out[COVT_BUF_A_COORDS], n_ints, xy)) != COVT_OK)
            die(ctx, "covt_result_read", rc);
        uint64_t h = 1469598103934665603ull; /* FNV-1a over the coordinate ints */
        if (L->status == COVT_OK)
            for (uint64_t i = 0; i < n_ints; i++) { h ^= (uint32_t)xy[i]; h *= 1099511628211ull; }
        free(xy);
        printf("%u %u %.*s %u %u %u %u %u %u %016llx\n", L->tile, L->layer_index, (int)L->name_length,
               (const char*)blob + L->name_offset, L->num_features, L->n_parts, L->n_rings, L->n_vertices, L->n_coords, L->status,
               (unsigned long long)h);
    }
    /* ... and the other direction (EncodingUtils.encodeByteRle, EncodingUtils.java:136): the decoded geometry_types stream of every layer,
     * encoded again on the GPU in ONE covt_encode_streams call, must give back the bytes of the tile */
    {
        uint64_t total = 0;
        uint32_t n_enc = 0;
        for (uint32_t l = 0; l < n_layers; l++)
            if (layers[l].status == COVT_OK) { total += (layers[l].num_features + 7u) & ~7ull; n_enc++; }
        uint8_t* values = (uint8_t*)calloc(total + 8, 1);
        covt_encode_desc* descs = (covt_encode_desc*)calloc(n_enc + 1, sizeof(covt_encode_desc));
        uint64_t at = 0;
        uint32_t k = 0;
        for (uint32_t l = 0; l < n_layers; l++) {
            const covt_layer* L = &layers[l];
            if (L->status != COVT_OK) continue;
            if (L->num_features && (rc = covt_result_read(res, COVT_BUF_S_GEOMETRY_TYPES, L->out[COVT_BUF_S_GEOMETRY_TYPES], L->num_features, values + at)) != COVT_OK)
                die(ctx, "covt_result_read", rc);
            descs[k].value_offset = at;
            descs[k].num_values = L->num_features;
            descs[k].op = COVT_OP_BYTE_RLE;
            at += (L->num_features + 7u) & ~7ull;
            k++;
        }
        covt_result* enc = NULL;
        if ((rc = covt_encode_streams(ctx, values, total, descs, n_enc, COVT_FLAG_DEFAULT, &enc)) != COVT_OK) die(ctx, "covt_encode_streams", rc);
        uint32_t same = 0;
        k = 0;
        for (uint32_t l = 0; l < n_layers; l++) {
            const covt_layer* L = &layers[l];
            if (L->status != COVT_OK) continue;
            const covt_stream_ref* S = &L->streams[COVT_SLOT_TYPES];
            uint8_t* bytes = (uint8_t*)malloc(descs[k].byte_length + 1);
            if (descs[k].status == COVT_OK && descs[k].byte_length &&
                (rc = covt_result_read(enc, COVT_BUF_STREAM_ARENA, descs[k].out_offset, descs[k].byte_length, bytes)) != COVT_OK)
                die(ctx, "covt_result_read", rc);
            same += descs[k].status == COVT_OK && descs[k].byte_length == S->byte_length && !memcmp(bytes, blob + S->byte_offset, S->byte_length);
            free(bytes);
            k++;
        }
        fprintf(stderr, "re-encoded %u geometry_types streams on the GPU: %u identical to the tile bytes\n", n_enc, same);
        covt_result_free(enc);
        free(descs);
        free(values);
    }
    covt_timing t;
    covt_result_timing(res, &t);
    fprintf(stderr, "%u tiles, %u layers: upload %.3f ms, decode %.3f ms (%u kernel launches), %.1f MB of stream payload, %llu vertices\n",
            n_tiles, n_layers, t.h2d_ms, t.decode_ms, t.kernel_launches, t.payload_bytes / 1e6, (unsigned long long)t.vertices);
    uint32_t bad = 0;
    for (uint32_t i = 0; i < n_tiles; i++) bad += tile_status[i] != COVT_OK;
    covt_result_free(res);
    covt_host_unregister(ctx, blob);
    covt_destroy(ctx);
    free(blob);
    free(offs);
    return bad ? 3 : 0;
}
