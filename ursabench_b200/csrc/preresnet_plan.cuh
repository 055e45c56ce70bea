// PreResNet (BasicBlock) layer plan shared by the CUDA-core and the tensor-core BMA forwards:
// where every tensor of one bank row lives (model.parameters() order, tests/golden/layouts.json) and where its
// packed inference form goes.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace ursa {

constexpr int kMaxLayers = 96;

struct PrepEntry {
    int type;          // 0 conv3x3 -> [ci][tap][co], 1 conv1x1 -> [ci][co], 2 bn fold, 3 copy,
                       // 4 conv3x3 -> K-major [co][tap*cin + ci] split into tf32 hi (dst) / lo (dst2) planes
                       // 5 conv3x3 (cin == cout == C) -> fused-stage layout [tap][ci/4][2C rows][4]: rows are
                       //   [hi(co); lo(co)] when buf == 0 and [lo(co); hi(co)] when buf == 1 (see bma_conv_fused.cuh)
                       // 7 conv3x3 (stride-2 transition, cout = 2 cin) -> K-major FP16-split rows [co][tap][hi(cin) | lo'(cin)]
                       //   halves (conv3x3s2_f16_kernel); occupies cin*cout*9 floats
                       // 6 conv3x3 (cin == cout == C) -> FP16-split fused-stage layout [tap][ci/8][2C rows][8 halves]:
                       //   rows [hi(co); lo'(co)], w = hi + lo' * 2^-11 (see bma_conv_fused16.cuh); occupies cin*cout*9 floats
    int cin, cout;     // conv: channels; bn: cout = channels; copy: cout = count
    int64_t src;       // offset in the bank row (conv weight / bn weight / copy source)
    int64_t src2;      // bn: offset of bias in the bank row
    int64_t buf;       // bn: offset of running_mean in the buffer row (running_var follows at +C)
    int64_t dst;       // offset in the packed row
    int64_t dst2;      // type 4: offset of the lo plane
};

struct PrepTable {
    int n;
    PrepEntry e[kMaxLayers];
};

static __global__ void __launch_bounds__(256) preresnet_prep_kernel(const PrepTable t, const float *__restrict__ bank,
                                                              int64_t ld_bank, const float *__restrict__ bufbank,
                                                              int64_t ld_buf, float *__restrict__ packed,
                                                              int64_t ld_packed) {
    const int s = blockIdx.y;
    const PrepEntry e = t.e[blockIdx.x];
    const float *row = bank + (int64_t)s * ld_bank;
    const float *brow = bufbank + (int64_t)s * ld_buf;
    float *dst = packed + (int64_t)s * ld_packed + e.dst;
    if (e.type == 0 || e.type == 1) {
        const int taps = e.type == 0 ? 9 : 1;
        const int total = e.cout * e.cin * taps;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {       // i over dst [ci][tap][co]
            const int co = i % e.cout;
            const int tap = (i / e.cout) % taps;
            const int ci = i / (e.cout * taps);
            dst[i] = row[e.src + ((int64_t)co * e.cin + ci) * taps + tap];   // PyTorch [co][ci][kh][kw]
        }
    } else if (e.type == 4) {
        float *dlo = packed + (int64_t)s * ld_packed + e.dst2;
        const int total = e.cout * e.cin * 9;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {       // i over dst [co][tap][ci]
            const int ci = i % e.cin;
            const int tap = (i / e.cin) % 9;
            const int co = i / (e.cin * 9);
            const float w = row[e.src + ((int64_t)co * e.cin + ci) * 9 + tap];
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(w));
            const float h = __uint_as_float(hb);
            uint32_t lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(w - h));
            dst[i] = h;
            dlo[i] = __uint_as_float(lb);
        }
    } else if (e.type == 5) {
        const int C = e.cout;
        const int total = 9 * (C / 4) * 2 * C * 4;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {       // i over dst [tap][j][n][e4]
            const int e4 = i & 3;
            const int n = (i >> 2) % (2 * C);
            const int j = (i / (8 * C)) % (C / 4);
            const int tap = i / (2 * C * C);
            const int co = n % C, ci = 4 * j + e4;
            const bool lo_part = ((n / C) != 0) != (e.buf != 0);
            const float w = row[e.src + ((int64_t)co * C + ci) * 9 + tap];
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(w));
            const float h = __uint_as_float(hb);
            uint32_t lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(w - h));
            dst[i] = lo_part ? __uint_as_float(lb) : h;
        }
    } else if (e.type == 6) {
        const int C = e.cout;
        __half *dh = reinterpret_cast<__half *>(dst);
        const int total = 9 * (C / 8) * 2 * C * 8;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {       // i over dst [tap][j][n][e8]
            const int e8 = i & 7;
            const int n = (i >> 3) % (2 * C);
            const int j = (i / (16 * C)) % (C / 8);
            const int tap = i / (2 * C * C);
            const int co = n % C, ci = 8 * j + e8;
            const float w = row[e.src + ((int64_t)co * C + ci) * 9 + tap];
            const __half h = __float2half_rn(w);
            dh[i] = n < C ? h : __float2half_rn((w - __half2float(h)) * 2048.f);
        }
    } else if (e.type == 10) {
        // 1x1 stride-2 shortcut of the FP16-split path: K-major rows [co][ [hi(16 ch) | lo'(16 ch)] x cin / 16 ] halves, the
        // layout of one tap of type 7 (it runs as a tenth K block of conv3x3s2_f16_kernel on the raw residual rows)
        __half *dh = reinterpret_cast<__half *>(dst);
        const int total = e.cout * 2 * e.cin;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int k = i % (2 * e.cin);
            const int co = i / (2 * e.cin);
            const int seg = k >> 4;
            const int ci = (seg >> 1) * 16 + (k & 15);
            const float w = row[e.src + (int64_t)co * e.cin + ci];          // PyTorch [co][ci][1][1]
            const __half h = __float2half_rn(w);
            dh[i] = (seg & 1) == 0 ? h : __float2half_rn((w - __half2float(h)) * 2048.f);
        }
    } else if (e.type == 9) {
        // network conv1 (3 -> 16) as a 16 -> 16 conv of the FP16-split stage kernel: type-6 layout with the 13 missing
        // input channels zero, so the stem is just the first conv of the stage-1 chain (its input planes carry the image)
        const int C = 16;
        __half *dh = reinterpret_cast<__half *>(dst);
        const int total = 9 * (C / 8) * 2 * C * 8;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {       // i over dst [tap][j][n][e8]
            const int e8 = i & 7;
            const int n = (i >> 3) % (2 * C);
            const int j = (i / (16 * C)) % (C / 8);
            const int tap = i / (2 * C * C);
            const int co = n % C, ci = 8 * j + e8;
            const float w = ci < 3 ? row[e.src + ((int64_t)co * 3 + ci) * 9 + tap] : 0.f;
            const __half h = __float2half_rn(w);
            dh[i] = n < C ? h : __float2half_rn((w - __half2float(h)) * 2048.f);
        }
    } else if (e.type == 7) {
        // stride-2 transition conv of the FP16-split path: K-major rows [co][tap][ [hi(16 ch) | lo'(16 ch)] x cin / 16 ] halves
        __half *dh = reinterpret_cast<__half *>(dst);
        const int total = e.cout * 9 * 2 * e.cin;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int k = i % (2 * e.cin);
            const int tap = (i / (2 * e.cin)) % 9;
            const int co = i / (18 * e.cin);
            const int seg = k >> 4;                                   // 16-half segments: (channel group, hi / lo')
            const int ci = (seg >> 1) * 16 + (k & 15);
            const float w = row[e.src + ((int64_t)co * e.cin + ci) * 9 + tap];
            const __half h = __float2half_rn(w);
            dh[i] = (seg & 1) == 0 ? h : __float2half_rn((w - __half2float(h)) * 2048.f);
        }
    } else if (e.type == 2) {
        for (int c = threadIdx.x; c < e.cout; c += blockDim.x) {
            const float mean = brow[e.buf + c], var = brow[e.buf + e.cout + c];
            const float a = row[e.src + c] / sqrtf(var + 1e-5f);     // alpha = weight * invstd
            dst[c] = a;
            dst[e.cout + c] = row[e.src2 + c] - mean * a;            // beta = bias - mean * alpha
        }
    } else {
        for (int i = threadIdx.x; i < e.cout; i += blockDim.x) dst[i] = row[e.src + i];
    }
}

struct NetPlan {
    int n_blocks;               // per stage
    int C;
    PrepTable table;
    int64_t packed_floats;
    int64_t conv1_w;
    int64_t conv1_w16;          // tc == 3: conv1 packed as a 16 -> 16 FP16-split stage conv (type 9), else -1
    struct Block { int64_t bn1, w1, bn2, w2, ds, w1_lo, w2_lo, ds16; } blocks[3][8];   // w*_lo: tensor-core plan only;
                                                                                       // ds16: type-10 shortcut (tc == 3)
    int64_t bn_final, fc;
    int64_t D, NB;              // expected bank / buffer row lengths
};

// tc == 0: 3x3 filters packed [ci][tap][co] for the CUDA-core kernels; tc == 1: K-major tf32 hi / lo planes (layer-by-
// layer tcgen05 kernels); tc == 2: fused-stage layout (type 5) for every 3x3 conv with cin == cout, type 4 for the two
// stride-2 transition convs; tc == 3: as 2 with the FP16-split layout (type 6)
static bool build_plan(int depth, int C, NetPlan &pl, int tc = 0) {
    if (depth >= 44 || depth < 8 || (depth - 2) % 6 != 0 || C < 1) return false;
    const int n = (depth - 2) / 6;
    if (n > 8) return false;
    pl.n_blocks = n;
    pl.C = C;
    PrepTable &t = pl.table;
    t.n = 0;
    int64_t src = 0, buf = 0, dst = 0;
    int64_t last_lo = -1;
    auto add_conv = [&](int type, int cin, int cout, int fused_mode = 0) {
        PrepEntry &e = t.e[t.n++];
        e.type = type; e.cin = cin; e.cout = cout; e.src = src; e.src2 = 0; e.buf = fused_mode; e.dst = dst; e.dst2 = 0;
        const int64_t cnt = (int64_t)cin * cout * (type == 1 ? 1 : 9);
        src += cnt;
        const int64_t d = dst;
        dst += type == 5 ? 2 * cnt : cnt;
        if (type == 4) {                 // hi plane at d, lo plane right after; both 16-byte aligned (cnt % 4 == 0)
            e.dst2 = dst;
            last_lo = dst;
            dst += cnt;
        }
        return d;
    };
    const int t3 = tc == 0 ? 0 : (tc == 3 ? 7 : 4);       // the two stride-2 transition convs of the tensor-core plans
    auto add_bn = [&](int c) {
        PrepEntry &e = t.e[t.n++];
        e.type = 2; e.cin = 0; e.cout = c; e.src = src; e.src2 = src + c; e.buf = buf; e.dst = dst; e.dst2 = 0;
        src += 2 * c;
        buf += 2 * c;          // running_mean, running_var (num_batches_tracked is int64 and not in the float bank)
        const int64_t d = dst;
        dst += 2 * c;
        return d;
    };
    // parameter order = model.parameters(): conv1, per block [bn1.w, bn1.b, conv1, bn2.w, bn2.b, conv2, (downsample)],
    // bn.w, bn.b, fc.w, fc.b   (tests/golden/layouts.json)
    pl.conv1_w = add_conv(0, 3, 16);
    const int widths[3] = {16, 32, 64};
    int inpl = 16;
    for (int st = 0; st < 3; ++st) {
        for (int b = 0; b < n; ++b) {
            const int w = widths[st];
            NetPlan::Block &B = pl.blocks[st][b];
            B.bn1 = add_bn(inpl);
            B.w1 = (tc >= 2 && inpl == w) ? add_conv(tc == 2 ? 5 : 6, inpl, w, 0) : add_conv(t3, inpl, w);
            B.w1_lo = last_lo;
            B.bn2 = add_bn(w);
            B.w2 = tc == 2 ? add_conv(5, w, w, 1) : (tc == 3 ? add_conv(6, w, w, 0) : add_conv(t3, w, w));
            B.w2_lo = last_lo;
            B.ds = (b == 0 && st > 0) ? add_conv(1, inpl, w) : -1;
            B.ds16 = -1;
            if (tc == 3 && B.ds >= 0) {                 // second packing of the same weights (src was advanced by add_conv)
                dst = (dst + 3) & ~(int64_t)3;
                PrepEntry &e = t.e[t.n++];
                e.type = 10; e.cin = inpl; e.cout = w; e.src = src - (int64_t)inpl * w; e.src2 = 0; e.buf = 0; e.dst = dst; e.dst2 = 0;
                B.ds16 = dst;
                dst += (int64_t)inpl * w;               // cout * 2 cin halves
            }
            inpl = w;
        }
    }
    pl.bn_final = add_bn(64);
    {
        PrepEntry &e = t.e[t.n++];
        e.type = 3; e.cin = 0; e.cout = C * 64 + C; e.src = src; e.src2 = 0; e.buf = 0; e.dst = dst; e.dst2 = 0;
        pl.fc = dst;
        src += C * 64 + C;
        dst += C * 64 + C;
    }
    pl.conv1_w16 = -1;
    if (tc == 3) {
        dst = (dst + 3) & ~(int64_t)3;          // bulk-copy source: 16-byte aligned
        PrepEntry &e = t.e[t.n++];
        e.type = 9; e.cin = 3; e.cout = 16; e.src = 0; e.src2 = 0; e.buf = 0; e.dst = dst; e.dst2 = 0;
        pl.conv1_w16 = dst;
        dst += 16 * 16 * 9;
    }
    pl.packed_floats = (dst + 3) & ~(int64_t)3;
    pl.D = src;
    pl.NB = buf;
    return t.n <= kMaxLayers;
}

}  // namespace ursa
