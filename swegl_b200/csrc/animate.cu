// animate.cu — scene_t::animate (swegl/data/model.hpp:146-177) and the node-hierarchy product of
// vertex_shader_t::original_to_world (swegl/render/vertex_shaders.hpp:16-33) on the device.
//
// The key frames, the nodes' base TRS and the hierarchy are uploaded once (swegl_b200_set_animation); a frame then
// carries only its time stamp (FrameParams::anim_time, 4 bytes in the frame block) and k_animate, the first kernel of
// the frame's chain, writes node_world / node_normal into the device frame block in front of k_vertex:
//
//   phase A  thread = node: start from the node's base TRS, apply its channels in channel order -- the two key frames
//            around fmod(t, end_time) (std::lower_bound on step.time < t, model.hpp:108-121) blended linearly in fp32
//            (model.hpp:156-163), rotation: quaternion normalised (vec2f.hpp:67-77) and matrix44_t::from_quaternion
//            (matrix44.hpp:34-43) -- then local = translate(scale(rotation, scale)) (model.hpp:55-60, points.cpp:14-30);
//   phase B  one pass per hierarchy level: world = parent_world * local with freon's product (s = 0; s += a[i][k] * b[k][j],
//            oracle/shims/freon/Matrix.hpp -- roots multiply the identity, which turns a -0 entry into +0 exactly as on the
//            host).
//
// Every operation is the reference's own fp32 operation in its order (no FMA: the TU is built with -fmad=false), so the
// matrices equal the host's bit for bit (tests/test_animation.py compares them with Scene.animate + node_matrices, which is
// pinned against the unmodified reference).  animate() in the reference edits the nodes in place, but every call rewrites
// the same (node, path) set from the time alone, so evaluating from the base TRS gives the state after ANY sequence of
// calls that ends with animate(t).
#include "common.cuh"

namespace sb {

// libstdc++'s std::lower_bound probe sequence on step.time < t (identical result on sorted keys; the same probes -- and
// therefore the same answer -- on unsorted ones and for a NaN t, for which every comparison is false)
SB_DEV uint32_t lower_bound_time(const float *time, uint32_t n, float t)
{
    uint32_t first = 0, len = n;
    while (len > 0) {
        const uint32_t half = len >> 1, mid = first + half;
        if (time[mid] < t) { first = mid + 1; len = len - half - 1; }
        else len = half;
    }
    return first;
}

// out = a * b, freon::operator* as the shim defines it
SB_DEV void matmul44(const float *a, const float *b, float *out)
{
    #pragma unroll
    for (int i = 0; i < 4; i++)
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            float s = 0.0f;
            #pragma unroll
            for (int k = 0; k < 4; k++) s = fadd(s, fmul(a[4 * i + k], b[4 * k + j]));
            out[4 * i + j] = s;
        }
}

__global__ void __launch_bounds__(ANIM_TPB) k_animate(AnimTables a, const FrameParams *__restrict__ fpp, float *node_world, float *node_normal)
{
    // (no pdl_trigger: the next kernel reads the matrices in its prologue, so it must not start before this one is done)
    const float t_in = fpp->anim_time;
    for (uint32_t i = threadIdx.x; i < a.n_nodes; i += ANIM_TPB) {
        float R[16], T[3], S[3];
        #pragma unroll
        for (int k = 0; k < 16; k++) R[k] = a.base_rotation[16 * i + k];
        #pragma unroll
        for (int k = 0; k < 3; k++) { T[k] = a.base_translation[3 * i + k]; S[k] = a.base_scale[3 * i + k]; }
        for (uint32_t c = a.node_chan_off[i]; c < a.node_chan_off[i + 1]; c++) {
            const AnimChannel ch = a.channels[c];
            const float rel = fmodf(t_in, ch.end_time);                     // model.hpp:150 (exact, like the host's fmod)
            const float *time = a.step_time + ch.first_step;
            const uint32_t it = lower_bound_time(time, ch.n_steps, rel);    // get_steps, model.hpp:108-121
            uint32_t b, af;
            if (it == ch.n_steps) b = af = ch.n_steps - 1;
            else if (it == 0) b = af = 0;
            else { b = it - 1; af = it; }
            const float tb = time[b], ta = time[af];
            const float4 vb = a.step_value[ch.first_step + b], va = a.step_value[ch.first_step + af];
            float4 f = vb;
            if (ta != tb) {                                                 // model.hpp:156-163
                const float wb = fdiv(fsub(ta, rel), fsub(ta, tb)), wa = fdiv(fsub(rel, tb), fsub(ta, tb));
                f.x = fadd(fmul(vb.x, wb), fmul(va.x, wa)); f.y = fadd(fmul(vb.y, wb), fmul(va.y, wa));
                f.z = fadd(fmul(vb.z, wb), fmul(va.z, wa)); f.w = fadd(fmul(vb.w, wb), fmul(va.w, wa));
            }
            if (ch.path == ANIM_PATH_ROTATION) {
                const float ln = __fsqrt_rn(fadd(fadd(fadd(fmul(f.x, f.x), fmul(f.y, f.y)), fmul(f.z, f.z)), fmul(f.w, f.w)));
                if (ln != 0.0f) { f.x = fdiv(f.x, ln); f.y = fdiv(f.y, ln); f.z = fdiv(f.z, ln); f.w = fdiv(f.w, ln); }
                const float q0 = f.x, q1 = f.y, q2 = f.z, q3 = f.w;         // from_quaternion, matrix44.hpp:34-43
                #pragma unroll
                for (int k = 0; k < 16; k++) R[k] = 0.0f;
                R[0] = fsub(fmul(2.0f, fadd(fmul(q0, q0), fmul(q1, q1))), 1.0f);
                R[1] = fmul(2.0f, fsub(fmul(q1, q2), fmul(q0, q3)));
                R[2] = fmul(2.0f, fadd(fmul(q1, q3), fmul(q0, q2)));
                R[4] = fmul(2.0f, fadd(fmul(q1, q2), fmul(q0, q3)));
                R[5] = fsub(fmul(2.0f, fadd(fmul(q0, q0), fmul(q2, q2))), 1.0f);
                R[6] = fmul(2.0f, fsub(fmul(q2, q3), fmul(q0, q1)));
                R[8] = fmul(2.0f, fsub(fmul(q1, q3), fmul(q0, q2)));
                R[9] = fmul(2.0f, fadd(fmul(q2, q3), fmul(q0, q1)));
                R[10] = fsub(fmul(2.0f, fadd(fmul(q0, q0), fmul(q3, q3))), 1.0f);
                R[15] = 1.0f;
            } else if (ch.path == ANIM_PATH_TRANSLATION) { T[0] = f.x; T[1] = f.y; T[2] = f.z; }
            else if (ch.path == ANIM_PATH_SCALE) { S[0] = f.x; S[1] = f.y; S[2] = f.z; }
        }
        // get_local_world_matrix, model.hpp:55-60: scale() multiplies columns 0..2 of all four rows, translate() adds to column 3
        #pragma unroll
        for (int r = 0; r < 4; r++)
            #pragma unroll
            for (int c = 0; c < 3; c++) R[4 * r + c] = fmul(R[4 * r + c], S[c]);
        #pragma unroll
        for (int r = 0; r < 3; r++)
            #pragma unroll
            for (int c = 0; c < 3; c++) node_normal[9 * i + 3 * r + c] = R[4 * r + c];
        #pragma unroll
        for (int r = 0; r < 3; r++) R[4 * r + 3] = fadd(R[4 * r + 3], T[r]);
        #pragma unroll
        for (int k = 0; k < 16; k++) a.local[16 * i + k] = R[k];
    }
    __syncthreads();
    // vertex_shaders.hpp:16-33, level by level (a node's parent is always one level up)
    for (uint32_t l = 0; l < a.n_levels; l++) {
        for (uint32_t k = a.level_off[l] + threadIdx.x; k < a.level_off[l + 1]; k += ANIM_TPB) {
            const uint32_t i = a.order[k];
            const int32_t p = a.parent[i];
            float P[16], L[16], W[16];
            #pragma unroll
            for (int e = 0; e < 16; e++) { L[e] = a.local[16 * i + e]; P[e] = p < 0 ? ((e % 5 == 0) ? 1.0f : 0.0f) : node_world[16 * p + e]; }
            matmul44(P, L, W);
            #pragma unroll
            for (int e = 0; e < 16; e++) node_world[16 * i + e] = W[e];
        }
        __syncthreads();
    }
}

void launch_animate(const AnimTables &a, const FrameParams *d_fp, float *node_world, float *node_normal, cudaStream_t st)
{
    if (a.n_nodes) k_animate<<<1, ANIM_TPB, 0, st>>>(a, d_fp, node_world, node_normal);
}

} // namespace sb
