// shim: the external "profi" profiler is not part of the reference repo; scopes compile to nothing.
#pragma once
#define PROFI_SCOPE(x)
#define PROFI_SCOPE_S2(x)
#define PROFI_SCOPE_S3(x)
#define PROFI_FUNC
