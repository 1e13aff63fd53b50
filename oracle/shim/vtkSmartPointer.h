// Stand-in for VTK's vtkSmartPointer (VTK is absent from this image).
// TEST INFRASTRUCTURE ONLY: lets the unmodified reference translation units
// (/root/reference/project/src/*.cpp) compile for the parity oracle. It covers
// exactly the members the reference touches (object3d_base.cpp:3-52,
// object2d.cpp:11-28): New(), operator->, implicit T* conversion, construction
// from a raw (non-owning) pointer.
#pragma once
#include <memory>

template <class T>
class vtkSmartPointer {
public:
    vtkSmartPointer() = default;
    vtkSmartPointer(T* borrowed) : _p(borrowed, [](T*) {}) {}
    static vtkSmartPointer New() {
        vtkSmartPointer s;
        s._p = std::make_shared<T>();
        return s;
    }
    T* operator->() const { return _p.get(); }
    T& operator*() const { return *_p; }
    operator T*() const { return _p.get(); }
    T* Get() const { return _p.get(); }

private:
    std::shared_ptr<T> _p;
};
