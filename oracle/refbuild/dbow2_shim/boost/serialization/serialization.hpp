// TEST INFRASTRUCTURE ONLY -- compile-only stand-in for the two boost.serialization names DBoW2's BowVector.h / FeatureVector.h
// mention (friend class boost::serialization::access; base_object<T>(*this) inside a serialize() template nobody instantiates).
#pragma once
namespace boost { namespace serialization {
class access {};
template <class Base, class Derived> Base& base_object(Derived& d) { return static_cast<Base&>(d); }
}}
