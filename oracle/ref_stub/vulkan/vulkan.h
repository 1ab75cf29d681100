/* ORACLE build helper (test infrastructure): opaque stand-ins for the few Vulkan types the reference's pipe headers mention,
 * so that the HOST side of its modules (crop/main.c, colour/main.c: roi and parameter arithmetic, no GPU work) compiles
 * in place for oracle/_ref.  not Vulkan, not shipped, never included by the product. */
#pragma once
#include <stdint.h>
typedef uint32_t VkFlags; typedef uint64_t VkDeviceSize; typedef uint32_t VkBool32;
#define VKH(N) typedef struct N##_T *N;
VKH(VkInstance) VKH(VkPhysicalDevice) VKH(VkDevice) VKH(VkQueue) VKH(VkSemaphore) VKH(VkCommandBuffer) VKH(VkFence) VKH(VkDeviceMemory)
VKH(VkBuffer) VKH(VkImage) VKH(VkEvent) VKH(VkQueryPool) VKH(VkBufferView) VKH(VkImageView) VKH(VkShaderModule) VKH(VkPipelineCache)
VKH(VkPipelineLayout) VKH(VkRenderPass) VKH(VkPipeline) VKH(VkDescriptorSetLayout) VKH(VkSampler) VKH(VkDescriptorPool) VKH(VkDescriptorSet)
VKH(VkFramebuffer) VKH(VkCommandPool) VKH(VkSurfaceKHR) VKH(VkSwapchainKHR) VKH(VkAccelerationStructureKHR) VKH(VkSamplerYcbcrConversion) VKH(VkDebugUtilsMessengerEXT)
typedef int VkResult; typedef int VkFormat; typedef int VkImageLayout; typedef int VkColorSpaceKHR; typedef int VkPresentModeKHR;
#define VK_SUCCESS 0
#define VK_INCOMPLETE 5
typedef VkFlags VkMemoryPropertyFlags; typedef VkFlags VkSubgroupFeatureFlags;
#define VK_MAX_PHYSICAL_DEVICE_NAME_SIZE 256
#define VK_MAX_MEMORY_TYPES 32
typedef struct { uint32_t memoryTypeCount; struct { VkMemoryPropertyFlags propertyFlags; uint32_t heapIndex; } memoryTypes[32]; uint32_t memoryHeapCount; struct { VkDeviceSize size; VkFlags flags; } memoryHeaps[16]; } VkPhysicalDeviceMemoryProperties;
typedef struct { char extensionName[256]; uint32_t specVersion; } VkExtensionProperties;
typedef struct { char layerName[256]; uint32_t specVersion, implementationVersion; char description[256]; } VkLayerProperties;
typedef void (*PFN_vkCmdPushDescriptorSetKHR)(void);
typedef struct { int dummy; } VkAccelerationStructureGeometryKHR;
typedef struct { int dummy; } VkAccelerationStructureBuildGeometryInfoKHR;
/* the module pass of the graph (graph-run-modules.h) builds a descriptor set layout and waits on a semaphore in between its
 * host-side work: the types and constants those lines mention */
#define VK_NULL_HANDLE 0
#define VK_STRUCTURE_TYPE_SEMAPHORE_WAIT_INFO 1
#define VK_STRUCTURE_TYPE_DESCRIPTOR_SET_LAYOUT_CREATE_INFO 2
#define VK_SHADER_STAGE_ALL 0x7fffffff
#define VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER 6
#define VK_DESCRIPTOR_SET_LAYOUT_CREATE_UPDATE_AFTER_BIND_POOL_BIT 2
typedef struct { uint32_t binding; int descriptorType; uint32_t descriptorCount; VkFlags stageFlags; const void *pImmutableSamplers; } VkDescriptorSetLayoutBinding;
typedef struct { int sType; const void *pNext; VkFlags flags; uint32_t bindingCount; const VkDescriptorSetLayoutBinding *pBindings; } VkDescriptorSetLayoutCreateInfo;
typedef struct { int sType; const void *pNext; VkFlags flags; uint32_t semaphoreCount; const VkSemaphore *pSemaphores; const uint64_t *pValues; } VkSemaphoreWaitInfo;
