/* gw_ctx.c -- the gateways' process-wide libpdegpu context. */
#include "gw_common.h"

static pdegpu_ctx *g_ctx = NULL;

static void gw_ctx_release(void)
{
    if (g_ctx) { pdegpu_free(g_ctx); g_ctx = NULL; }
}

pdegpu_ctx *gw_ctx(const char *gw)
{
    if (!g_ctx) {
        const char *env = getenv("PDEGPU_DEVICE");
        int dev = env ? atoi(env) : 0;
        int rc = pdegpu_init(dev, &g_ctx);
        if (rc != PDEGPU_OK) {
            char msg[600];
            snprintf(msg, sizeof(msg), "%s: %s", gw, pdegpu_last_error(NULL));
            g_ctx = NULL;
            mexErrMsgTxt(msg);
        }
        mexAtExit(gw_ctx_release);
    }
    /* the sweep order follows the environment at EVERY call: setenv('PDEGPU_ORDER', 'reference') in a running Matlab /
     * Octave session switches the unchanged drivers to the reference's line order (include/pdegpu.h) */
    pdegpu_set_sweep_order(g_ctx, pdegpu_order_from_env());
    return g_ctx;
}
