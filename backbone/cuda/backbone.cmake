#########################################################################
# BBM CUDA Backbone (B200, sm_100a) - drop this directory next to backbone/native
# and configure with -DBBM_BACKBONE=cuda (cmake/bbm_helpers.cmake:21-42 includes
# backbone/<name>/backbone.cmake; backbone/native/backbone.cmake:8,13 is the model).
#
# Host-side scalar code keeps the native backbone's types (a single eval on the host
# is the reference's own code); the batched / fitting paths go to libbbmcu.so through
# the adapters in include/bbm_cuda/*.h.  BBMCU_ROOT = checkout of the bbm-b200 repository.
#########################################################################

#########################################################################
# Set available configurations
#########################################################################
set(BBM_BACKBONE_CONFIGURATIONS "floatRGB" "doubleRGB" "floatRGB_cuda")

#########################################################################
# Set Include dirs: this backbone first (its backbone.h wraps the native one)
#########################################################################
if(NOT DEFINED BBMCU_ROOT)
  get_filename_component(BBMCU_ROOT "${CMAKE_CURRENT_LIST_DIR}/../.." ABSOLUTE)
endif()
target_include_directories(${BBM_NAME} INTERFACE ${CMAKE_CURRENT_LIST_DIR}/include)
target_include_directories(${BBM_NAME} INTERFACE ${BBM_SOURCE_DIR}/backbone/native/include)
target_include_directories(${BBM_NAME} INTERFACE ${BBMCU_ROOT}/include)

#########################################################################
# Link the C ABI of the CUDA kernels
#########################################################################
find_library(BBMCU_LIBRARY NAMES bbmcu PATHS ${BBMCU_ROOT}/bbm_b200 NO_DEFAULT_PATH REQUIRED)
target_link_libraries(${BBM_NAME} INTERFACE ${BBMCU_LIBRARY})
