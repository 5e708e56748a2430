/* Shim for include/gfdm/api.h:25.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_GR_ATTRIBUTES_H
#define ORACLE_SHIM_GR_ATTRIBUTES_H
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
#endif
