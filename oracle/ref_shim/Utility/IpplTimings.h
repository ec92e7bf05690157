// stand-in for Utility/IpplTimings.h (oracle/ref_shim): timers are no-ops.
#pragma once
#include <string>
class IpplTimings {
public:
    using TimerRef = unsigned;
    static TimerRef getTimer(const char*) { return 0; }
    static void startTimer(TimerRef) {}
    static void stopTimer(TimerRef) {}
};
