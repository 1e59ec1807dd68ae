// oracle/ref_thrust_stub.cpp -- TEST INFRASTRUCTURE ONLY.  libgpvref_refcuda.so links the reference's host objects against the
// reference's own kernels (cuda/CUDAClassifyTessellation.cu); the one other device symbol those objects reference lives in
// cuda/THRUSTUtilities.cu (Thrust), which is not built here and is never called on the voxelizer path.
#include <cstdlib>
extern "C" float THRUSTDeviceFindMax(float*, int, int) { abort(); }
