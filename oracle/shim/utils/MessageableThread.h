// Stand-in (empty) for glare-core utils/MessageableThread.h; AudioResampler.h includes it but uses nothing from it.
#pragma once
