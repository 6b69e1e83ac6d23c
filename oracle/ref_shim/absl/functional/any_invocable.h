// Test-infrastructure shim (NOT product code): the hnswlib headers include this but use nothing from it.
#ifndef VK_ORACLE_SHIM_ABSL_ANY_INVOCABLE_H_
#define VK_ORACLE_SHIM_ABSL_ANY_INVOCABLE_H_
#endif
