#!/bin/bash
# On the GPU box: compute-sanitizer memcheck and racecheck over a spread of parity scenes (small, large, clipped,
# overflow lists and the big-list fallback, big-triangle queue, overlays, device NewMesh) and over the round-2 paths:
# host mirrors + the one-call Draw (CUDA graph), batches split by the workspace limit, strip draws with the reject
# pre-pass and the list-walking setup kernel.  Both tools must report 0 errors / 0 hazards.
set -x
K1="c1_suzanne_800x600 or wire_verts_multi_object or c2_cube_poseB or stacked_quads or tiny_far_dense or odd_size or new_mesh_on_device or fog_wire"
K2="c1_serial_tiles1 or wire_c1 or tiny_far or big_triangles or gouraud_textured or stacked_quads or c2_cube_poseA or inside_sphere or multi_object"
K3="object_moves or option_changes or batch_mirror or strips_and_odd or workspace_limit or overflow_and_big or strips_compose"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -q -k "$K1" 2>&1 | tail -4
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_mirror_gpu.py tests/test_parity_gpu.py -q -k "$K3" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -q -k "$K2" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_mirror_gpu.py tests/test_parity_gpu.py -q -k "$K3" 2>&1 | tail -4
