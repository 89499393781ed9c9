// xsb_ghost.cu -- ghost operators -- placeholder, filled in below.
#include "xsb_ctx.h"
void xsb_ghost_release(xsb_ctx*) {}
extern "C" {
int xsb_comm_unique_id(void*) { return XSB_ERR_UNSUPPORTED; }
int xsb_comm_init(xsb_ctx* ctx, int, int, const void*) { return ctx ? ctx->fail(XSB_ERR_UNSUPPORTED, "not implemented") : XSB_ERR_STATE; }
int xsb_ghost_comm_scheme(xsb_ctx* ctx, const xsb_domain_desc*, const uint64_t*) { return ctx ? ctx->fail(XSB_ERR_UNSUPPORTED, "not implemented") : XSB_ERR_STATE; }
int xsb_ghost_update(xsb_ctx* ctx, uint32_t) { return ctx ? ctx->fail(XSB_ERR_UNSUPPORTED, "not implemented") : XSB_ERR_STATE; }
int xsb_ghost_reduce_add(xsb_ctx* ctx, uint32_t) { return ctx ? ctx->fail(XSB_ERR_UNSUPPORTED, "not implemented") : XSB_ERR_STATE; }
}
