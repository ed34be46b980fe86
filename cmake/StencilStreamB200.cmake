# StencilStream-B200 — CMake entry point of the B200 backend.
#
# The reference selects a backend by linking an INTERFACE target (reference CMakeLists.txt:39-51:
# `StencilStream_CUDA` adds `-fsycl -fsycl-targets=nvptx64-nvidia-cuda --offload-arch=sm_80` and defines
# STENCILSTREAM_BACKEND_CUDA / STENCILSTREAM_TARGET_CUDA; per example e.g.
# examples/hotspot/CMakeLists.txt:19-25, examples/jacobi/CMakeLists.txt:20). This module provides the
# counterpart for this backend,
#
#     include(<repo>/cmake/StencilStreamB200.cmake)          # instead of the reference's root project
#     add_executable(hotspot_cuda hotspot.cpp)
#     target_link_libraries(hotspot_cuda PUBLIC StencilStream_B200)   # was: StencilStream_CUDA
#     target_compile_definitions(hotspot_cuda PUBLIC HOTSPOT_SPLIT_CELL_STRUCT=1)
#
# i.e. the link target is the only line of an example's CMake file that changes. What the target
# stands for:
#   * the same two macros, so that the examples' `#if STENCILSTREAM_BACKEND_*` ladders pick their cuda
#     branch, and the include roots of this backend (`StencilStream/...`, the `sycl::` vocabulary shim,
#     the C ABI headers) instead of the reference's root and the SYCL headers;
#   * nvcc for sm_100a, C++20, --expt-relaxed-constexpr, -lineinfo; libstst_rt (the C-ABI runtime);
#   * the source annotator (stencilstream_b200/tools/annotate.py): nvcc has no device-by-default mode
#     and the reference's functors carry no `__host__ __device__`, so every
#     executable that links StencilStream_B200 is compiled from a build-tree copy of its source
#     directory in which functions taking a `Stencil<...> const &` are prefixed with STST_HD. The
#     sources under version control stay untouched. This happens in a deferred finalizer at the end of
#     the top-level CMakeLists.txt (cmake_language(DEFER), CMake >= 3.19), at configure time; the
#     original files are registered as configure dependencies, so editing them re-runs it.
#
# `stencilstream_b200_add_executable(<name> <sources...> [ALSO fn1 fn2 ...])` is the explicit form
# (ALSO names further functions device code calls that do not take the stencil, e.g. a cell's
# `halo()` factory); `set_property(TARGET <name> PROPERTY STST_B200_ALSO fn1 fn2)` does the same for a
# target declared with plain add_executable.
#
# STST_B200_REFERENCE_TARGETS=ON (set before including this file) additionally defines the reference's
# target names: `StencilStream_CUDA` = link StencilStream_B200, the others (StencilStream_CPU, the FPGA
# ones) as empty placeholders. With it the reference's example directories configure with their own,
# unmodified CMakeLists.txt (tests/cmake_dropin/CMakeLists.txt add_subdirectory()s them) and their
# `*_cuda` targets build against this backend.
#
# Options: STST_B200_FMAD (ON; OFF passes -fmad=false: no a*b+c contraction, results bit-identical to
# the reference's cpu backend built with -ffp-contract=off), STST_B200_PYTHON (interpreter that runs
# the annotator).

cmake_minimum_required(VERSION 3.24)
include_guard(GLOBAL)

get_filename_component(STST_B200_ROOT "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
set(STST_B200_ROOT "${STST_B200_ROOT}" CACHE INTERNAL "StencilStream-B200 repository root")

option(STST_B200_FMAD "Let nvcc contract a*b+c into FMAs (nvcc's default)" ON)
if(NOT STST_B200_PYTHON)
    find_program(STST_B200_PYTHON NAMES python3 python REQUIRED)
endif()

if(NOT DEFINED CMAKE_CUDA_ARCHITECTURES)
    set(CMAKE_CUDA_ARCHITECTURES 100a)
endif()
enable_language(CUDA)

# ---- the C-ABI runtime (include/stst_rt.h) -------------------------------------------------------------
if(NOT TARGET stst_rt)
    add_library(stst_rt SHARED "${STST_B200_ROOT}/stencilstream_b200/csrc/stst_rt.cu")
    target_include_directories(stst_rt PUBLIC "${STST_B200_ROOT}/include")
    set_target_properties(stst_rt PROPERTIES CUDA_STANDARD 20 CUDA_STANDARD_REQUIRED ON
                                             CUDA_ARCHITECTURES 100a POSITION_INDEPENDENT_CODE ON)
    target_compile_options(stst_rt PRIVATE $<$<COMPILE_LANGUAGE:CUDA>:-lineinfo>)
    target_link_libraries(stst_rt PRIVATE ${CMAKE_DL_LIBS})
endif()

# ---- the backend target ----------------------------------------------------------------------------------
add_library(StencilStream_B200 INTERFACE)
target_include_directories(StencilStream_B200 INTERFACE
    "${STST_B200_ROOT}/stencilstream_b200/include"
    "${STST_B200_ROOT}/stencilstream_b200/compat"
    "${STST_B200_ROOT}/include")
target_compile_definitions(StencilStream_B200 INTERFACE STENCILSTREAM_BACKEND_CUDA=1
                                                        STENCILSTREAM_TARGET_CUDA=1)
target_compile_features(StencilStream_B200 INTERFACE cuda_std_20)
target_compile_options(StencilStream_B200 INTERFACE
    $<$<COMPILE_LANGUAGE:CUDA>:--expt-relaxed-constexpr -lineinfo -w>
    $<$<AND:$<COMPILE_LANGUAGE:CUDA>,$<NOT:$<BOOL:${STST_B200_FMAD}>>>:-fmad=false>)
target_link_libraries(StencilStream_B200 INTERFACE stst_rt)

# The reference's other switch (CMakeLists.txt:62-63); kernels of this backend are templates named by
# their functor anyway, the macro is accepted for source compatibility.
if(NOT TARGET StencilStream_NamedKernels)
    add_library(StencilStream_NamedKernels INTERFACE)
    target_compile_definitions(StencilStream_NamedKernels INTERFACE STENCILSTREAM_NAMED_KERNELS)
endif()

if(STST_B200_REFERENCE_TARGETS)
    add_library(StencilStream_CUDA INTERFACE)
    target_link_libraries(StencilStream_CUDA INTERFACE StencilStream_B200)
    # Placeholders: this repository provides the cuda backend only. Targets that link one of these
    # still configure (the examples' CMake files declare all their variants) but are not buildable.
    foreach(other Base CPU FPGABase VerboseSynthesis MonotileBase MonotileEmulator Monotile
                  MonotileReport TilingBase TilingEmulator Tiling TilingReport)
        if(NOT TARGET StencilStream_${other})
            add_library(StencilStream_${other} INTERFACE)
        endif()
    endforeach()
endif()

define_property(TARGET PROPERTY STST_B200_ALSO
    BRIEF_DOCS "Further functions the annotator makes device-callable"
    FULL_DOCS "Names of functions defined in the target's sources that device code calls although they do not take the stencil")

function(stencilstream_b200_add_executable name)
    cmake_parse_arguments(ARG "" "" "ALSO" ${ARGN})
    add_executable(${name} ${ARG_UNPARSED_ARGUMENTS})
    target_link_libraries(${name} PUBLIC StencilStream_B200)
    if(ARG_ALSO)
        set_property(TARGET ${name} PROPERTY STST_B200_ALSO ${ARG_ALSO})
    endif()
endfunction()

# ---- finalizer: annotated build-tree copies, compiled as CUDA ------------------------------------------
function(_stst_b200_collect_targets dir out)
    get_property(targets DIRECTORY "${dir}" PROPERTY BUILDSYSTEM_TARGETS)
    get_property(subdirs DIRECTORY "${dir}" PROPERTY SUBDIRECTORIES)
    foreach(sub IN LISTS subdirs)
        _stst_b200_collect_targets("${sub}" sub_targets)
        list(APPEND targets ${sub_targets})
    endforeach()
    set(${out} ${targets} PARENT_SCOPE)
endfunction()

function(_stst_b200_finalize)
    _stst_b200_collect_targets("${CMAKE_SOURCE_DIR}" all_targets)
    foreach(tgt IN LISTS all_targets)
        get_target_property(type ${tgt} TYPE)
        if(NOT type STREQUAL "EXECUTABLE")
            continue()
        endif()
        get_target_property(libs ${tgt} LINK_LIBRARIES)
        if(NOT libs)
            continue()
        endif()
        if(NOT "StencilStream_B200" IN_LIST libs AND
           NOT (STST_B200_REFERENCE_TARGETS AND "StencilStream_CUDA" IN_LIST libs))
            continue()
        endif()
        get_target_property(src_dir ${tgt} SOURCE_DIR)
        get_target_property(bin_dir ${tgt} BINARY_DIR)
        get_target_property(sources ${tgt} SOURCES)
        get_target_property(also ${tgt} STST_B200_ALSO)
        set(also_arg "")
        set(also_csv "")
        if(also)
            list(JOIN also "," also_csv)
            set(also_arg "--also=${also_csv}")
        endif()
        # one annotated copy per (source directory, ALSO list), shared by the targets built from it
        string(MD5 stage_key "${src_dir}|${also_csv}")
        string(SUBSTRING "${stage_key}" 0 8 stage_key)
        set(stage "${bin_dir}/b200_src_${stage_key}")
        get_property(staged_already GLOBAL PROPERTY _STST_B200_STAGED_${stage_key})
        if(NOT staged_already)
        set_property(GLOBAL PROPERTY _STST_B200_STAGED_${stage_key} ON)
        execute_process(
            COMMAND "${CMAKE_COMMAND}" -E env "PYTHONPATH=${STST_B200_ROOT}"
                    "${STST_B200_PYTHON}" -m stencilstream_b200.tools.annotate --tree "${src_dir}"
                    -o "${stage}" ${also_arg}
            RESULT_VARIABLE rc OUTPUT_VARIABLE log ERROR_VARIABLE log)
        if(NOT rc EQUAL 0)
            message(FATAL_ERROR "StencilStream-B200: annotating ${src_dir} failed:\n${log}")
        endif()
        file(GLOB_RECURSE originals "${src_dir}/*.cpp" "${src_dir}/*.hpp" "${src_dir}/*.h")
        set_property(DIRECTORY "${CMAKE_SOURCE_DIR}" APPEND PROPERTY CMAKE_CONFIGURE_DEPENDS ${originals})
        endif()
        set(staged "")
        set(include_dirs "${stage}")
        foreach(src IN LISTS sources)
            if(NOT IS_ABSOLUTE "${src}")
                set(src "${src_dir}/${src}")
            endif()
            file(RELATIVE_PATH rel "${src_dir}" "${src}")
            if(rel MATCHES "^\\.\\.")
                message(FATAL_ERROR "StencilStream-B200: source ${src} of ${tgt} lies outside ${src_dir}")
            endif()
            list(APPEND staged "${stage}/${rel}")
            get_filename_component(rel_dir "${stage}/${rel}" DIRECTORY)
            list(APPEND include_dirs "${rel_dir}")
        endforeach()
        list(REMOVE_DUPLICATES include_dirs)
        set_property(TARGET ${tgt} PROPERTY SOURCES ${staged})
        set_source_files_properties(${staged} TARGET_DIRECTORY ${tgt} PROPERTIES LANGUAGE CUDA)
        target_include_directories(${tgt} PRIVATE ${include_dirs})
        set_target_properties(${tgt} PROPERTIES CUDA_STANDARD 20 CUDA_STANDARD_REQUIRED ON
                                                CUDA_ARCHITECTURES 100a LINKER_LANGUAGE CUDA)
        message(STATUS "StencilStream-B200: ${tgt} builds from annotated copies in ${stage}")
    endforeach()
endfunction()

cmake_language(DEFER DIRECTORY "${CMAKE_SOURCE_DIR}" CALL _stst_b200_finalize)
