// test shim: csrc/velo_functors.h compiled for the host (g++), exported for ctypes — tests/test_functors_host.py
#include "velo_functors.h"
extern "C" int functor_eval(int type, const double *k, const double *pose, double *r, double *J) {
    RotPack P, Pi;
    rotpack_make(pose, false, &P);
    rotpack_make(pose, true, &Pi);
    switch (type) {
    case 0: lin3d3d(k, P, pose + 3, r, J); return 3;
    case 1: lin3d2d(k, P, pose + 3, r, J); return 2;
    case 2: lin2d3d(k, Pi, pose + 3, r, J); return 2;
    case 3: lin2d2d(k, P, pose + 3, r, J); return 1;
    case 4: lin3dpd(k, P, pose + 3, r, J); return 1;
    }
    return -1;
}
