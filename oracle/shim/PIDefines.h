/* oracle/shim/PIDefines.h -- TEST INFRASTRUCTURE ONLY: calling-convention macros, empty off Windows/Mac. */
#ifndef ORACLE_SHIM_PIDEFINES_H
#define ORACLE_SHIM_PIDEFINES_H
#define DLLExport
#define MACPASCAL
#endif
