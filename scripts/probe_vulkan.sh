#!/bin/bash
# can the reference (vkdt-cli: Vulkan + glslang) run on this box?  evidence for DESIGN.md section 5 / BASELINE.md:
# loader, ICDs (NVIDIA's and Mesa lavapipe), tools, headers.  writes to stdout; the caller keeps it under profiles/.
echo "== date / host"; date -u; uname -a; nproc
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,driver_version --format=csv,noheader 2>&1 | head -8
echo "== vulkaninfo"; (command -v vulkaninfo && vulkaninfo --summary) 2>&1 | head -40 || echo "vulkaninfo: not found"
echo "== loader libraries (ldconfig -p | grep -i vulkan)"; ldconfig -p | grep -i -E "vulkan|libGLX_nvidia|libnvidia-vulkan|lvp|libEGL_nvidia" || echo "none"
echo "== find libvulkan / nvidia vulkan producer / lavapipe"; find / -xdev \( -name 'libvulkan*' -o -name 'libnvidia-vulkan-producer*' -o -name 'libvulkan_lvp*' -o -name 'libGLX_nvidia*' -o -name 'libnvidia-glcore*' -o -name 'libnvidia-gpucomp*' \) 2>/dev/null | head -20; echo "(end of find)"
echo "== ICD json files"; ls -la /usr/share/vulkan/icd.d /etc/vulkan/icd.d /usr/local/share/vulkan/icd.d 2>&1 | head -20
echo "== shader compilers"; for t in glslangValidator glslc glslang spirv-opt; do printf "%s: " $t; command -v $t || echo "not found"; done
echo "== headers"; ls /usr/include/vulkan/vulkan.h /usr/local/include/vulkan/vulkan.h 2>&1
echo "== toolchains for the un-vendored decoders"; for t in cargo rustc clang cmake; do printf "%s: " $t; command -v $t || echo "not found"; done
echo "== VK env"; env | grep -i -E "^VK_|VULKAN" || echo "none"
echo "== python vulkan bindings"; python -c "import vulkan" 2>&1 | tail -1
