// Headless stand-in for <SFML/Graphics.hpp> -- test infrastructure only.
// The reference renders through SFML (GraphicPrinter.hpp); the oracle harness never
// opens a window, so every class here is an inert shell with the members that
// GraphicPrinter.hpp / basic.hpp name.  Nothing here is shipped in the product.
#pragma once
#include <string>

namespace sf {

struct Color {
    unsigned char r = 0, g = 0, b = 0;
    Color() {}
    Color(int r_, int g_, int b_) : r(r_), g(g_), b(b_) {}
    static const Color White, Black, Red, Green, Yellow, Blue, Magenta, Cyan;
};
inline const Color Color::White(255, 255, 255);
inline const Color Color::Black(0, 0, 0);
inline const Color Color::Red(255, 0, 0);
inline const Color Color::Green(0, 255, 0);
inline const Color Color::Yellow(255, 255, 0);
inline const Color Color::Blue(0, 0, 255);
inline const Color Color::Magenta(255, 0, 255);
inline const Color Color::Cyan(0, 255, 255);

struct Vector2f {
    float x = 0, y = 0;
    Vector2f() {}
    Vector2f(float x_, float y_) : x(x_), y(y_) {}
};

struct FloatRect { float left = 0, top = 0, width = 0, height = 0; };

struct Drawable {};

struct ConvexShape : Drawable {
    explicit ConvexShape(int = 0) {}
    void setFillColor(const Color &) {}
    void setPoint(int, Vector2f) {}
    void setPosition(float, float) {}
};

struct RectangleShape : Drawable {
    explicit RectangleShape(Vector2f = Vector2f()) {}
    void setFillColor(const Color &) {}
    void setPosition(float, float) {}
};

struct Font {
    bool loadFromFile(const std::string &) { return true; }
};

struct Text : Drawable {
    Text(const std::string &, const Font &, unsigned int) {}
    void setFillColor(const Color &) {}
    void setPosition(float, float) {}
    FloatRect getLocalBounds() const { return FloatRect(); }
};

struct Event {
    enum EventType { Closed, Other };
    EventType type = Other;
};

struct VideoMode {
    VideoMode(unsigned int, unsigned int) {}
};

struct RenderWindow {
    RenderWindow(VideoMode, const std::string &) {}
    void setFramerateLimit(unsigned int) {}
    bool isOpen() const { return false; }
    bool pollEvent(Event &) { return false; }
    void close() {}
    void clear(const Color &) {}
    void draw(const Drawable &) {}
    void display() {}
};

struct Time { long long us = 0; };
inline Time microseconds(long long v) { Time t; t.us = v; return t; }
inline void sleep(Time) {}

} // namespace sf
