/* Process exit codes of the drivers -- the reference's contract (inc/enums.h:4-11):
 * 0 success, 1 device problems, 2 program/queue/buffer/launch problems, 3 file problems, 4 other. */
#ifndef B200_HOST_ENUMS_H
#define B200_HOST_ENUMS_H

typedef enum {
    Success = 0,
    OpenCLDeviceError = 1,  /* name kept: here it means "no usable CUDA device" */
    OpenCLProgramError = 2, /* name kept: any b200_* runtime/launch failure */
    FileError = 3,
    OtherError = 4
} ReturnCode;

#endif
