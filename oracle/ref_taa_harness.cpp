// oracle/ref_taa_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or shipped with the product).
//
// The reference's OWN temporal-AA pass, PassTemporalAAAdapter (shs/pipeline/pass_adapters.hpp:1402-1491),
// compiled where it lies under /root/reference.  pass_adapters.hpp pulls in the Jolt-guarded culling headers; they compile against
// the JoltPhysics declaration shim oracle/jolt_shim (the TAA pass itself touches no Jolt type).  Same C signature as the
// restatement's shso_pass_taa.  Built by oracle/Makefile (`make ref`) into oracle/_ref/libshs_taa_ref.so.
#include <cstdint>
#include <cstring>

#define SHS_HAS_JOLT 1
#include "shs/core/context.hpp"
#include "shs/pipeline/pass_adapters.hpp"

extern "C" int32_t shsref_pass_taa(uint8_t* ldr_inout, uint8_t* history_inout, int32_t history_valid, int32_t w, int32_t h)
{
    if (!ldr_inout || !history_inout || w <= 0 || h <= 0) return 1;
    const size_t n = (size_t)w * (size_t)h;
    static_assert(sizeof(shs::Color) == 4, "RGBA8");
    shs::RT_ColorLDR ldr(w, h);
    std::memcpy(ldr.color.data.data(), ldr_inout, n * 4);
    shs::Context ctx{};
    if (history_valid)
    {
        ctx.temporal_aa.history.resize(n);
        std::memcpy(ctx.temporal_aa.history.data(), history_inout, n * 4);
        ctx.temporal_aa.history_w = w;
        ctx.temporal_aa.history_h = h;
        ctx.temporal_aa.history_valid = true;
    }
    shs::RTRegistry rtr{};
    shs::PassTemporalAAAdapter pass(rtr.reg<shs::RTHandle>(&ldr));
    shs::PassExecutionRequest req{};
    req.inputs.registry = &rtr;
    req.valid = true;
    pass.execute_resolved(ctx, req);
    std::memcpy(ldr_inout, ldr.color.data.data(), n * 4);
    if (ctx.temporal_aa.history.size() == n) std::memcpy(history_inout, ctx.temporal_aa.history.data(), n * 4);
    return ctx.temporal_aa.history_valid ? 0 : 2;
}
