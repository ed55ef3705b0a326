/* pdegpu_flow_batch -- batch command-line front end of libpdegpu's device-resident flow drivers.
 *
 *   pdegpu_flow_batch [--driver llin|fmg|hs] [--device N] [--batch B] ROWS COLS CHANNELS LIST.txt
 *
 * LIST.txt names one frame pair per line: "frame0.raw frame1.raw out_prefix". A .raw file is ROWS x COLS x CHANNELS
 * single-precision values in Matlab's column-major order (what fwrite(fid, single(I), 'single') produces), values
 * 0..255. For every pair the flow is written to out_prefix_U.raw / out_prefix_V.raw (ROWS x COLS single, same order).
 * Pairs are processed B at a time by ONE call of pdegpu_flow_llin_2d / pdegpu_flow_fmg_2d / pdegpu_flow_hs_2d
 * (the whole FlowEminND_llin_2D_v10 / FlowEminNDFASFMG_elin_2D_v10 / FlowEminHS_elin_2D_v10 driver with its default
 * parameters, matlab/optical_flow/). Data-parallel over several GPUs: start one process per GPU with --device and a
 * share of the list (pairs are independent).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "pdegpu.h"

static int read_raw(const char *path, float *dst, size_t n)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "pdegpu_flow_batch: cannot open %s\n", path); return -1; }
    const size_t got = fread(dst, sizeof(float), n, f);
    fclose(f);
    if (got != n) { fprintf(stderr, "pdegpu_flow_batch: %s holds %zu of %zu values\n", path, got, n); return -1; }
    return 0;
}

static int write_raw(const char *prefix, const char *suffix, const float *src, size_t n)
{
    char path[4096];
    snprintf(path, sizeof path, "%s%s", prefix, suffix);
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "pdegpu_flow_batch: cannot create %s\n", path); return -1; }
    const size_t put = fwrite(src, sizeof(float), n, f);
    fclose(f);
    return put == n ? 0 : -1;
}

int main(int argc, char **argv)
{
    const char *driver = "llin";
    int device = 0, batch = 16, a = 1;
    for (; a < argc && argv[a][0] == '-' && argv[a][1] == '-'; a++) {
        if (!strcmp(argv[a], "--driver") && a + 1 < argc) driver = argv[++a];
        else if (!strcmp(argv[a], "--device") && a + 1 < argc) device = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--batch") && a + 1 < argc) batch = atoi(argv[++a]);
        else { fprintf(stderr, "unknown option %s\n", argv[a]); return 2; }
    }
    if (argc - a != 4 || batch < 1) {
        fprintf(stderr, "usage: pdegpu_flow_batch [--driver llin|fmg|hs] [--device N] [--batch B] ROWS COLS CHANNELS LIST.txt\n");
        return 2;
    }
    const int rows = atoi(argv[a]), cols = atoi(argv[a + 1]), ch = atoi(argv[a + 2]);
    FILE *list = fopen(argv[a + 3], "r");
    if (!list || rows < 8 || cols < 8 || ch < 1) { fprintf(stderr, "pdegpu_flow_batch: bad arguments\n"); return 2; }

    pdegpu_ctx *ctx = NULL;
    if (pdegpu_init(device, &ctx) != PDEGPU_OK) { fprintf(stderr, "pdegpu_init: %s\n", pdegpu_last_error(NULL)); return 1; }
    pdegpu_flow_llin_params pl; pdegpu_flow_fmg_params pf; pdegpu_flow_hs_params ph;
    pdegpu_flow_llin_default_params(&pl); pdegpu_flow_fmg_default_params(&pf); pdegpu_flow_hs_default_params(&ph);

    const size_t nimg = (size_t)rows * cols * ch, nflow = (size_t)rows * cols;
    float *I0 = malloc(nimg * batch * sizeof(float)), *I1 = malloc(nimg * batch * sizeof(float));
    float *U = malloc(nflow * batch * sizeof(float)), *V = malloc(nflow * batch * sizeof(float));
    char (*prefix)[4096] = malloc((size_t)batch * 4096);
    if (!I0 || !I1 || !U || !V || !prefix) { fprintf(stderr, "pdegpu_flow_batch: out of memory\n"); return 1; }

    char f0[4096], f1[4096];
    long done = 0;
    int rc = 0, eof = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    while (!eof && !rc) {
        int nb = 0;
        while (nb < batch) {
            if (fscanf(list, "%4095s %4095s %4095s", f0, f1, prefix[nb]) != 3) { eof = 1; break; }
            if (read_raw(f0, I0 + nimg * nb, nimg) || read_raw(f1, I1 + nimg * nb, nimg)) { rc = 1; break; }
            nb++;
        }
        if (rc || nb == 0) break;
        if (!strcmp(driver, "llin")) rc = pdegpu_flow_llin_2d(ctx, U, V, I0, I1, rows, cols, ch, nb, &pl);
        else if (!strcmp(driver, "fmg")) rc = pdegpu_flow_fmg_2d(ctx, U, V, I0, I1, rows, cols, ch, nb, &pf);
        else if (!strcmp(driver, "hs")) rc = pdegpu_flow_hs_2d(ctx, U, V, I0, I1, rows, cols, ch, nb, &ph);
        else { fprintf(stderr, "unknown driver %s\n", driver); rc = 2; break; }
        if (rc != PDEGPU_OK) { fprintf(stderr, "libpdegpu: %s\n", pdegpu_last_error(ctx)); break; }
        for (int b = 0; b < nb && !rc; b++)
            rc = write_raw(prefix[b], "_U.raw", U + nflow * b, nflow) || write_raw(prefix[b], "_V.raw", V + nflow * b, nflow);
        done += nb;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double s = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    fprintf(stderr, "pdegpu_flow_batch: %ld pairs in %.3f s (%.1f flows/s incl. file I/O), driver %s, GPU %d\n", done, s, done / (s > 0 ? s : 1), driver, device);
    fclose(list);
    free(I0); free(I1); free(U); free(V); free(prefix);
    pdegpu_free(ctx);
    return rc ? 1 : 0;
}
