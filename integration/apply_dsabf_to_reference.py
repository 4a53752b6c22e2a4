#!/usr/bin/env python3
"""Apply the INTEGRATION.md drop-in to a scratch copy of the reference's own main().

    apply_dsabf_to_reference.py <reference>/src  <out_dir>

Copies the reference's translation unit (beamformer.cu and the headers it includes) to <out_dir> and makes the four
edits a maintainer would make to run the hot path through libdsabf.so while keeping everything else of main() --
option parsing, weight loop, test generator / PSRDADA reader, observation_loop_state, the H2D copies into the
reference's own d_data ring, its streams, the D2H copies, the DEBUG cublasSgemv and bin/data.py:

  A  #include "dsabf.h"
  B  one handle, created next to the reference's other set-up, fed the reference's own fourier_coefficients[]
  C  expand_input<<<>>> + cublasGemmStridedBatchedEx + detect_sum<<<>>>   (src/beamformer.cu:464-481)
        ->  ONE dsabf_beamform_device() on the same stream, same input chunk, same d_out slot
  D  dsabf_destroy() in the teardown

The edits are anchored on the statements themselves (not on line numbers) and the script fails if an anchor is
missing or ambiguous.  Nothing of the reference is stored in this repository: the copy and a unified diff of it
(reference_dsabf.patch) are written to <out_dir> only.  oracle/Makefile builds <out_dir>/beamformer.cu into
oracle/_ref/beam_ref_dsabf; tests/test_host_programs.py checks its bin/data.py against the unpatched reference's.
"""
import difflib
import os
import shutil
import sys


def find_one(lines, needle, start=0):
    hits = [i for i in range(start, len(lines)) if needle in lines[i]]
    if len(hits) != 1:
        sys.exit("apply_dsabf_to_reference: expected exactly one line containing %r, found %d" % (needle, len(hits)))
    return hits[0]


def main():
    src_dir, out_dir = sys.argv[1], sys.argv[2]
    os.makedirs(out_dir, exist_ok=True)
    for name in os.listdir(src_dir):
        if name.endswith((".hh", ".cuh", ".cu", ".h")):
            shutil.copy(os.path.join(src_dir, name), os.path.join(out_dir, name))
    path = os.path.join(out_dir, "beamformer.cu")
    orig = open(path).read().splitlines(keepends=True)
    L = list(orig)

    # D (bottom-up so that earlier indices stay valid)
    i = find_one(L, '"Freeing CUDA Structures"')
    L.insert(i, "\tgpuErrchk((cudaError_t) dsabf_destroy(bf));\n")

    # C: the three launches of one chunk -> one fused call
    a = find_one(L, "expand_input<<<")
    b = find_one(L, "detect_sum<<<", a)
    g = find_one(L, "cublasGemmStridedBatchedEx(", a)
    if not (a < g < b) or not L[b].rstrip().endswith(");"):
        sys.exit("apply_dsabf_to_reference: the expand/GEMM/detect sequence does not look as expected")
    while a > 0 and L[a - 1].strip().startswith("/*") and "Expand" in L[a - 1]:
        a -= 1
    fused = [
        "\t\t\t\t\t/* unpack + beamform + detect of this chunk: one fused sm_100a kernel (libdsabf.so) on this stream */\n",
        "\t\t\t\t\tgpuErrchk((cudaError_t) dsabf_beamform_device(bf,\n",
        "\t\t\t\t\t\t&d_data[N_BYTES_PRE_EXPANSION_PER_GEMM*(N_GEMMS_PER_BLOCK*obs_state.get_next_gpu_analysis_block() + timeSlice[st])],\n",
        "\t\t\t\t\t\t1, &d_out[st*N_F_PER_DETECT], (void*) stream[st]));\n",
    ]
    L[a:b + 1] = fused

    # B: the handle, right before the loop state is constructed
    i = find_one(L, "observation_loop_state obs_state(")
    setup = [
        "\t/* dsabf drop-in: the reference's shape constants are the library's defaults */\n",
        "\tdsabf_config bf_cfg;\n",
        "\t#if DEBUG\n", "\t\tdsabf_default_config(&bf_cfg, 1);\n", "\t#else\n", "\t\tdsabf_default_config(&bf_cfg, 0);\n", "\t#endif\n",
        "\tdsabf_handle* bf = NULL;\n",
        "\tgpuErrchk((cudaError_t) dsabf_create(&bf_cfg, &bf));\n",
        "\tgpuErrchk((cudaError_t) dsabf_set_weights(bf, (const int8_t*) fourier_coefficients));\n\n",
    ]
    L[i:i] = setup

    # A
    i = find_one(L, '#include "beamformer.cuh"')
    L.insert(i + 1, '#include "dsabf.h"\t\t// C ABI of the B200 fused path\n')

    open(path, "w").write("".join(L))
    patch = difflib.unified_diff(orig, L, "a/src/beamformer.cu", "b/src/beamformer.cu", n=2)
    open(os.path.join(out_dir, "reference_dsabf.patch"), "w").write("".join(patch))
    print("apply_dsabf_to_reference: wrote %s (+ reference_dsabf.patch)" % path)


if __name__ == "__main__":
    main()
