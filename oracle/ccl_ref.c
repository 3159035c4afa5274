/*
 * ccl_ref.c - plain-C CPU oracle (TEST INFRASTRUCTURE, never shipped/linked
 * into the product library).
 *
 * Restates what the reference obtains from third-party code on the hot path:
 *
 *  - cc3d.connected_components(bin_img, return_N=True)   count_blobs.py:61,64
 *    (connected-components-3d 3.12.3, requirements.txt:9, NOT vendored):
 *    26-connected labelling of the non-zero voxels of a C-order (Z,Y,X)
 *    volume; final labels 1..N numbered in the order in which each
 *    component's first voxel is met in a C-order raster scan.
 *  - cc3d.statistics(labels, no_slice_conversion=True)   count_blobs.py:85:
 *    per label 0..N voxel count, inclusive bounding box, centroid =
 *    sum(coord)/count (we return the exact integer sums; the fp64 divide is
 *    done by the caller so it is bit-identical to a double division).
 *  - scipy.ndimage.binary_erosion(mask, iterations=30, border_value=1)
 *    inference/inference.py:82: iterated erosion with the default 3-D cross
 *    (6-neighbour) structuring element, outside-of-array treated as 1.
 *
 * Algorithms are the textbook ones (two-pass union-find; literal iterated
 * erosion) - deliberately different from the GPU kernels they check.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static uint32_t uf_find(uint32_t *p, uint32_t a) {
    uint32_t r = a;
    while (p[r] != r) r = p[r];
    while (p[a] != r) { uint32_t n = p[a]; p[a] = r; a = n; }
    return r;
}

static uint32_t uf_union(uint32_t *p, uint32_t a, uint32_t b) {
    a = uf_find(p, a); b = uf_find(p, b);
    if (a == b) return a;
    if (a < b) { p[b] = a; return a; }
    p[a] = b; return b;
}

/* returns N (number of components) or -1 on allocation failure / overflow */
int64_t dlvref_ccl26(const uint8_t *mask, int64_t Z, int64_t Y, int64_t X, uint32_t *labels) {
    int64_t nvox = Z * Y * X;
    if (nvox == 0) return 0;
    /* worst case for 26-connectivity: one provisional label per 2x2x2 block + slack */
    int64_t cap = ((Z + 1) / 2) * ((Y + 1) / 2) * ((X + 1) / 2) + 2;
    if (cap > 0xFFFFFFF0LL) return -1;
    uint32_t *par = (uint32_t *)malloc((size_t)cap * sizeof(uint32_t));
    if (!par) return -1;
    uint32_t next = 1;
    par[0] = 0;
    for (int64_t z = 0; z < Z; ++z)
        for (int64_t y = 0; y < Y; ++y)
            for (int64_t x = 0; x < X; ++x) {
                int64_t i = (z * Y + y) * X + x;
                if (!mask[i]) { labels[i] = 0; continue; }
                uint32_t cur = 0;
                /* the 13 raster-scan predecessors of a voxel */
                for (int dz = -1; dz <= 0; ++dz)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            if (dz == 0 && (dy > 0 || (dy == 0 && dx >= 0))) continue;
                            int64_t zz = z + dz, yy = y + dy, xx = x + dx;
                            if (zz < 0 || yy < 0 || yy >= Y || xx < 0 || xx >= X) continue;
                            uint32_t l = labels[(zz * Y + yy) * X + xx];
                            if (!l) continue;
                            cur = cur ? uf_union(par, cur, l) : uf_find(par, l);
                        }
                if (!cur) {
                    if ((int64_t)next >= cap) { free(par); return -1; }
                    cur = next; par[next] = next; ++next;
                }
                labels[i] = cur;
            }
    /* roots are the minimum provisional label of their set (union by min) and
     * provisional labels are issued in raster order => ascending root order is
     * first-voxel order. */
    uint32_t *remap = (uint32_t *)calloc((size_t)next, sizeof(uint32_t));
    if (!remap) { free(par); return -1; }
    uint32_t n = 0;
    for (uint32_t l = 1; l < next; ++l)
        if (par[l] == l) remap[l] = ++n;
    for (uint32_t l = 1; l < next; ++l)
        if (par[l] != l) remap[l] = remap[uf_find(par, l)];
    for (int64_t i = 0; i < nvox; ++i) labels[i] = remap[labels[i]];
    free(remap);
    free(par);
    return (int64_t)n;
}

/* counts[N+1], sums[(N+1)*3] (z,y,x), bbox[(N+1)*6] = zmin,zmax,ymin,ymax,xmin,xmax (inclusive).
 * Labels absent from the volume keep count 0 and bbox (dim, -1). */
void dlvref_stats(const uint32_t *labels, int64_t Z, int64_t Y, int64_t X, int64_t N,
                  uint64_t *counts, uint64_t *sums, int64_t *bbox) {
    for (int64_t l = 0; l <= N; ++l) {
        counts[l] = 0;
        sums[3 * l] = sums[3 * l + 1] = sums[3 * l + 2] = 0;
        bbox[6 * l + 0] = Z; bbox[6 * l + 1] = -1;
        bbox[6 * l + 2] = Y; bbox[6 * l + 3] = -1;
        bbox[6 * l + 4] = X; bbox[6 * l + 5] = -1;
    }
    for (int64_t z = 0; z < Z; ++z)
        for (int64_t y = 0; y < Y; ++y)
            for (int64_t x = 0; x < X; ++x) {
                uint32_t l = labels[(z * Y + y) * X + x];
                counts[l]++;
                sums[3 * l] += (uint64_t)z; sums[3 * l + 1] += (uint64_t)y; sums[3 * l + 2] += (uint64_t)x;
                int64_t *b = bbox + 6 * l;
                if (z < b[0]) b[0] = z;
                if (z > b[1]) b[1] = z;
                if (y < b[2]) b[2] = y;
                if (y > b[3]) b[3] = y;
                if (x < b[4]) b[4] = x;
                if (x > b[5]) b[5] = x;
            }
}

/* literal iterated 6-neighbour erosion, border_value = 1 */
int dlvref_erode6(const uint8_t *mask, int64_t Z, int64_t Y, int64_t X, int iterations, uint8_t *out) {
    int64_t n = Z * Y * X;
    uint8_t *a = (uint8_t *)malloc((size_t)(n ? n : 1));
    if (!a) return -1;
    for (int64_t i = 0; i < n; ++i) out[i] = mask[i] ? 1 : 0;
    for (int it = 0; it < iterations; ++it) {
        memcpy(a, out, (size_t)n);
        int changed = 0;
        for (int64_t z = 0; z < Z; ++z)
            for (int64_t y = 0; y < Y; ++y)
                for (int64_t x = 0; x < X; ++x) {
                    int64_t i = (z * Y + y) * X + x;
                    if (!a[i]) continue;
                    int keep = 1;
                    if (x > 0 && !a[i - 1]) keep = 0;
                    else if (x < X - 1 && !a[i + 1]) keep = 0;
                    else if (y > 0 && !a[i - X]) keep = 0;
                    else if (y < Y - 1 && !a[i + X]) keep = 0;
                    else if (z > 0 && !a[i - Y * X]) keep = 0;
                    else if (z < Z - 1 && !a[i + Y * X]) keep = 0;
                    if (!keep) { out[i] = 0; changed = 1; }
                }
        if (!changed) break;
    }
    free(a);
    return 0;
}
