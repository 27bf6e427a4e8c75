/*
 * oracle/shim/lcms2.h -- TEST INFRASTRUCTURE ONLY.
 * Little-CMS is not in this image and the ICC step is outside the accelerated path:
 * the oracle always runs with "no transform".  These inert declarations only let ScopedLcms.h and
 * ColorProfileConversion.h (class layout) parse.
 */
#ifndef ORACLE_SHIM_LCMS2_H
#define ORACLE_SHIM_LCMS2_H
#include <stdint.h>
typedef void* cmsContext;
typedef void* cmsHPROFILE;
typedef void* cmsHTRANSFORM;
typedef struct _cms_MLU_struct cmsMLU;
typedef struct _cms_curve_struct cmsToneCurve;
typedef uint32_t cmsUInt32Number;
static inline void cmsMLUfree(cmsMLU*) {}
static inline void cmsFreeToneCurve(cmsToneCurve*) {}
static inline void cmsDeleteContext(cmsContext) {}
static inline int cmsCloseProfile(cmsHPROFILE) { return 1; }
static inline void cmsDeleteTransform(cmsHTRANSFORM) {}
#endif
