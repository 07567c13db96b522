/* reference-build shim: boost::mutex / boost::mutex::scoped_lock over the standard library (the only boost names the
 * map classes use: include/KeyFrame.h, MapPoint.h, Map.h, KeyFrameDatabase.h) */
#ifndef ORB_REFSHIM_BOOST_THREAD_HPP
#define ORB_REFSHIM_BOOST_THREAD_HPP
#include <mutex>
namespace boost {
class mutex {
public:
    typedef std::unique_lock<std::mutex> scoped_lock_base;
    class scoped_lock {
    public:
        explicit scoped_lock(mutex& m) : l(m.m) {}
    private:
        std::unique_lock<std::mutex> l;
    };
    mutex() {}
    mutex(const mutex&) {}                       /* KeyFrame / MapPoint objects are never copied with a held lock */
    mutex& operator=(const mutex&) { return *this; }
private:
    std::mutex m;
};
class thread { };
}
#endif
