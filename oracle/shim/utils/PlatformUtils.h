// Stand-in (empty) for glare-core utils/PlatformUtils.h; AudioResampler.cpp includes it but uses nothing from it.
#pragma once
