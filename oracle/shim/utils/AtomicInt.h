// Stand-in (empty) for glare-core utils/AtomicInt.h; AudioResampler.h includes it but uses nothing from it.
#pragma once
