// shim: the external "profi" profiler is not part of the reference repo.
#pragma once
