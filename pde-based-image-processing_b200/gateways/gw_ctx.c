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
    return g_ctx;
}
