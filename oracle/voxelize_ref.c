/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of hard voxelisation (point -> pillar grouping).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this file.  The product path (lidar_vision_vqa_b200/) never does.
 *
 * What it restates.  The reference calls a THIRD-PARTY voxeliser that is not vendored:
 *   src/lidar-encoder/pcdet/datasets/processor/data_processor.py:16-61  (VoxelGeneratorWrapper)
 *   src/lidar-encoder/pcdet/datasets/processor/data_processor.py:133-180 (transform_points_to_voxels)
 * which resolves to spconv `VoxelGeneratorV2` / `VoxelGenerator` (spconv 1.x) or
 * `Point2VoxelCPU3d.point_to_voxel` (spconv 2.x).  The reference does not pin a version
 * (docs/INSTALL.md allows 1.0 / 1.2 / 2.x; docker images install spconv-cu102 / spconv-cu116).
 * This file restates the published algorithm shared by spconv >=1.2 and 2.x
 * (`points_to_voxel_3d_np` / `Point2VoxelCPU::point_to_voxel`):
 *
 *   grid[j]  = round((range[3+j] - range[j]) / vsize[j])            (done by the caller, see
 *                                                                     data_processor.py:135-136)
 *   for i in 0..N-1, in point order:
 *     for j in x,y,z:  c = floor((p[i][j] - range[j]) / vsize[j])   in IEEE fp32, true division
 *                      if c < 0 or c >= grid[j]: skip the point
 *     coor = (cz, cy, cx)
 *     id = table[coor]
 *     if id == -1:
 *        if num_voxels >= max_voxels: skip the point (`continue`; existing voxels still accept points)
 *        id = num_voxels++ ; table[coor] = id ; coors[id] = coor
 *     if cnt[id] < max_points: voxels[id][cnt[id]] = p[i] ; cnt[id]++
 *
 * The in-tree statements it is cross-checked against (tests/test_oracle.py):
 *   quantisation formula   src/lidar-encoder/pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py:93-96
 *   (z,y,x) coordinate use src/lidar-encoder/pcdet/models/backbones_3d/vfe/pillar_vfe.py:100-103
 *                          src/lidar-encoder/pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:27
 * PARITY STATUS: the pillar SET and per-pillar counts are pinned against the reference's own
 * DynamicPillarVFE run in the build container (tests/golden/); the first-appearance id order and the
 * first-P cap follow spconv's published loop and are NOT pinned by anything inside /root/reference
 * ("parity unpinned" for the ordering -- spconv is absent and the reference ships no golden vectors).
 *
 * Non-finite coordinates: spconv stores floor() into an int (UB for NaN/Inf, INT_MIN on x86 => the
 * point is rejected).  Restated here as "reject unless 0 <= c < grid", which rejects NaN/Inf too.
 *
 * Build:  gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Returns the number of voxels produced for this frame (<= max_voxels), or -1 on bad arguments.
 *
 * points      [n, c] row-major fp32, columns 0..2 = x,y,z
 * range6      xmin,ymin,zmin,xmax,ymax,zmax  (fp32, as spconv stores them)
 * vsize3      voxel size x,y,z (fp32)
 * grid3       grid size x,y,z
 * voxels      [max_voxels, max_points, c]  written (zero-filled first)        -- may be NULL
 * coords      [max_voxels, 3] (z,y,x)                                          -- required
 * num_points  [max_voxels]                                                     -- required
 * point_voxel [n]  voxel id of each point, -1 if rejected/dropped              -- may be NULL
 * point_slot  [n]  slot inside the voxel, -1 if not stored (over the cap)      -- may be NULL
 */
int oracle_voxelize_hard(const float *points, int64_t n, int c, const float *range6, const float *vsize3,
                         const int32_t *grid3, int max_points, int max_voxels, float *voxels, int32_t *coords,
                         int32_t *num_points, int32_t *point_voxel, int32_t *point_slot)
{
    if (!points || !range6 || !vsize3 || !grid3 || !coords || !num_points || c < 3 || max_points < 1 ||
        max_voxels < 0)
        return -1;
    const int64_t gx = grid3[0], gy = grid3[1], gz = grid3[2];
    const int64_t cells = gx * gy * gz;
    int32_t *table = (int32_t *)malloc((size_t)cells * sizeof(int32_t));
    if (!table)
        return -1;
    memset(table, 0xFF, (size_t)cells * sizeof(int32_t)); /* -1 */
    if (voxels)
        memset(voxels, 0, (size_t)max_voxels * max_points * c * sizeof(float));
    memset(num_points, 0, (size_t)max_voxels * sizeof(int32_t));

    int num_voxels = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = points + i * c;
        int32_t cc[3];
        int ok = 1;
        for (int j = 0; j < 3; ++j) {
            volatile float d = p[j] - range6[j];      /* volatile: keep the two roundings separate */
            volatile float q = d / vsize3[j];
            float f = floorf(q);
            if (!(f >= 0.0f && f < (float)grid3[j])) {
                ok = 0;
                break;
            }
            cc[j] = (int32_t)f;
        }
        if (point_voxel)
            point_voxel[i] = -1;
        if (point_slot)
            point_slot[i] = -1;
        if (!ok)
            continue;
        const int64_t cell = ((int64_t)cc[2] * gy + cc[1]) * gx + cc[0];
        int32_t id = table[cell];
        if (id < 0) {
            if (num_voxels >= max_voxels)
                continue;
            id = num_voxels++;
            table[cell] = id;
            coords[3 * id + 0] = cc[2];
            coords[3 * id + 1] = cc[1];
            coords[3 * id + 2] = cc[0];
        }
        if (point_voxel)
            point_voxel[i] = id;
        const int32_t k = num_points[id];
        if (k < max_points) {
            if (voxels)
                memcpy(voxels + ((size_t)id * max_points + k) * c, p, (size_t)c * sizeof(float));
            if (point_slot)
                point_slot[i] = k;
            num_points[id] = k + 1;
        }
    }
    free(table);
    return num_voxels;
}

/* The reference's dense BEV scatter, restated for one batch:
 *   src/lidar-encoder/pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:14-37  (nz == 1)
 *   canvas[b][f][z + y*nx + x] = feats[m][f]   for every pillar m with coords (b, z, y, x); zero elsewhere;
 *   pointpillar_scatter.py:40-73 (PointPillarScatter3d, nz > 1): canvas[b][f][z*ny*nx + y*nx + x], the canvas
 *   [F, nz*ny*nx] then viewed as [F*nz, ny, nx].
 * coords is [m,4] int32 (b,z,y,x); bev is [B, F*nz, ny, nx] and is zero-filled here.  Later rows overwrite
 * earlier ones on a duplicate cell (the voxeliser never produces duplicates).
 */
int oracle_scatter_bev(const float *feats, const int32_t *coords, int64_t m, int nb, int f, int nx, int ny, int nz,
                       float *bev)
{
    if (!feats || !coords || !bev || nz < 1)
        return -1;
    const size_t plane = (size_t)nx * ny * nz;
    memset(bev, 0, (size_t)nb * f * plane * sizeof(float));
    for (int64_t i = 0; i < m; ++i) {
        const int32_t b = coords[4 * i], z = coords[4 * i + 1], y = coords[4 * i + 2], x = coords[4 * i + 3];
        if (b < 0 || b >= nb || y < 0 || y >= ny || x < 0 || x >= nx || z < 0 || z >= nz)
            return -2;
        const size_t cell = nz > 1 ? ((size_t)z * ny + y) * nx + x           /* pointpillar_scatter.py:63 */
                                   : (size_t)z + (size_t)y * nx + x;         /* pointpillar_scatter.py:27 */
        float *dst = bev + (size_t)b * f * plane + cell;
        const float *src = feats + (size_t)i * f;
        for (int k = 0; k < f; ++k)
            dst[(size_t)k * plane] = src[k];
    }
    return 0;
}
