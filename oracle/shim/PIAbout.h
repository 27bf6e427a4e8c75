/* oracle/shim/PIAbout.h -- TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_PIABOUT_H
#define ORACLE_SHIM_PIABOUT_H
typedef struct AboutRecord AboutRecord;
typedef AboutRecord* AboutRecordPtr;
#endif
