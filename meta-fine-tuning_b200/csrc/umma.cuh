// tcgen05 / TMEM / mbarrier primitives (filled in with the tensor-core path).
#pragma once
