/* <gsl/gsl_histogram.h> is included by the reference (RVNode.h:8) but every use is
 * commented out (RVNode.cpp:35-53,68-69).  Empty on purpose. */
#ifndef GBN_SHIM_GSL_HISTOGRAM_H
#define GBN_SHIM_GSL_HISTOGRAM_H
#endif
